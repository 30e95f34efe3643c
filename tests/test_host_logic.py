# -*- coding: utf-8 -*-
"""
CPU tests of the host logic: tiling (mirror of the reference's
tests/test_core.py:36-67), the on-disk memmap layout, Xrange host helpers, the
Xrange scalar arithmetic of the oracle (mirror of tests/test_numba_xr.py), the
C-ABI surface (every symbol of include/*.h is exported) and the loud failure
of the product path when no GPU / no library is present.
"""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest

import oracle_lib as ol
import fractalshades_b200 as fsb
import fractalshades_b200.models as fsm
from fractalshades_b200 import _native, settings

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nx,xy_ratio", [
    (800, 1.0), (600, 1.5), (3840, 16 / 9.), (200, 1.0), (201, 1.0), (199, 0.7),
    (401, 2.0), (1000, 1.8), (123, 1.23), (7680, 16 / 9.), (64, 1.0)])
def test_chunk_indexing_consistent(nx, xy_ratio):
    """ chunk_slices / chunks_count / chunk_rank / chunk_from_rank agree
    (reference tests/test_core.py:36-67) """
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=0., y=0., dx=1., nx=nx, xy_ratio=xy_ratio, theta_deg=0.)
    slices = list(f.chunk_slices())
    assert len(slices) == f.chunks_count
    tot = 0
    for i, cs in enumerate(slices):
        assert f.chunk_rank(cs) == i
        assert f.chunk_from_rank(i) == cs
        tot += (cs[1] - cs[0]) * (cs[3] - cs[2])
    assert tot == f.nx * f.ny


def test_known_tile_counts():
    """ 800^2 -> 16 tiles, 4K -> 220, 8K -> 858 """
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    for nx, ratio, n in ((800, 1.0, 16), (3840, 16 / 9., 220), (7680, 16 / 9., 858)):
        f.zoom(x=0., y=0., dx=1., nx=nx, xy_ratio=ratio, theta_deg=0.)
        assert f.chunks_count == n


def test_pixel_grid_layout():
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=0., y=0., dx=1., nx=400, xy_ratio=2.0, theta_deg=0.)
    p = f.chunk_pixel_pos((0, 200, 0, 200), False, None)
    assert p.shape == (200, 200)
    assert p[0, 0].real == -0.5 and p[0, 0].imag == pytest.approx(0.25)
    assert p[0, 1].real > p[0, 0].real and p[1, 0].imag < p[0, 0].imag
    pj = f.chunk_pixel_pos((0, 200, 0, 200), 1.0, 2)
    assert pj.shape == (400, 400)


def test_xrange_host_helpers():
    import mpmath
    from fractalshades_b200 import xrange as fsx
    mpmath.mp.dps = 1100
    m, e = fsx.mpf_to_xr(mpmath.mpf("1e-1000"))
    assert 0.5 <= m < 1.0 and abs(float(mpmath.ldexp(m, e) / mpmath.mpf("1e-1000")) - 1) < 1e-15
    m, e = fsx.mpc_to_xr(mpmath.mpc("3e-500", "-4e-510"))
    assert abs(m.real) >= 0.5 and e < -1600
    m, e = fsx.xr_complex_from_parts(0.75, -1100, 0.5, -1105)
    assert (m, e) == (complex(0.75, 0.5 / 32.), -1100)
    assert fsx.xr_complex_from_parts(0.0, 0, 0.5, -30) == (complex(0., 0.5), -30)


def _rand_xr(seed, n, complex_=True):
    """ like generate_random_xr of the reference's tests/test_numba_xr.py:23-46 """
    rg = np.random.default_rng(seed)
    m = (rg.random(n) * 2. - 1.) * np.exp2(rg.integers(-60, 60, n).astype(float))
    if complex_:
        m = m + 1j * (rg.random(n) * 2. - 1.) * np.exp2(rg.integers(-60, 60, n).astype(float))
    e = rg.integers(-2000, 2000, n).astype(np.int32)
    return m, e


def _to_mp(m, e):
    import mpmath
    mpmath.mp.prec = 200
    if np.iscomplexobj(m):
        return [mpmath.mpc(mpmath.ldexp(float(a.real), int(b)), mpmath.ldexp(float(a.imag), int(b)))
                for a, b in zip(m, e)]
    return [mpmath.ldexp(float(a), int(b)) for a, b in zip(m, e)]


@pytest.mark.parametrize("op", [0, 1, 2])
def test_oracle_xr_binop_against_mpmath(op):
    """ add / sub / mul of Xrange complex scalars, checked in 200-bit
    arithmetic (reference: tests/test_numba_xr.py:303-421) """
    import mpmath
    a, ae = _rand_xr(100, 300)
    b, be = _rand_xr(800, 300)
    if op < 2:       # comparable magnitudes so that the sum is not trivial
        be = (ae + np.random.default_rng(5).integers(-3, 3, ae.size)).astype(np.int32)
    out, oe = ol.xr_binop_c(op, a, ae, b, be)
    A, B, O = _to_mp(a, ae), _to_mp(b, be), _to_mp(out, oe)
    for x, y, z in zip(A, B, O):
        ref = x + y if op == 0 else (x - y if op == 1 else x * y)
        scale = max(abs(x), abs(y)) if op < 2 else abs(ref)
        assert abs(z - ref) <= scale * mpmath.mpf(2) ** -50


def test_oracle_xr_to_standard_and_normalize():
    a, ae = _rand_xr(7800, 200)
    ae = (ae // 4).astype(np.int32)              # within 2^+-800
    std = ol.xr_to_standard_c(a, ae)
    exp = np.ldexp(a.real, ae) + 1j * np.ldexp(a.imag, ae)
    ok = np.isclose(std, exp, rtol=4e-16, atol=0) | ((np.abs(exp) < 1e-290))
    assert ok.all()
    n, ne = ol.xr_normalize_c(a, ae)
    big = np.maximum(np.abs(n.real), np.abs(n.imag))
    assert np.all((big >= 1.0) & (big < 2.0))     # numba_xr.py:674-685: [1, 2)
    assert np.allclose(ol.xr_to_standard_c(n, ne), std, rtol=1e-15, atol=0)


def test_oracle_xr_real_ops_and_compare():
    a, ae = _rand_xr(1, 200, complex_=False)
    b, be = _rand_xr(2, 200, complex_=False)
    be = (ae + np.random.default_rng(3).integers(-2, 2, ae.size)).astype(np.int32)
    A, B = _to_mp(a, ae), _to_mp(b, be)
    for cmp, fn in enumerate((lambda x, y: x < y, lambda x, y: x <= y,
                              lambda x, y: x == y, lambda x, y: x != y,
                              lambda x, y: x >= y, lambda x, y: x > y)):
        got = ol.xr_compare_f(cmp, a, ae, b, be)
        assert list(got) == [bool(fn(x, y)) for x, y in zip(A, B)]
    # Xrange_scalar(1.0, 10) == 1024. (tests/test_numba_xr.py:774-778)
    assert ol.xr_to_standard_f(np.array([1.0]), np.array([10], np.int32))[0] == 1024.


def test_hypot_definition_within_one_ulp():
    rg = np.random.default_rng(0)
    x = rg.standard_normal(2000) * np.exp2(rg.integers(-1000, 1000, 2000).astype(float))
    y = rg.standard_normal(2000) * np.exp2(rg.integers(-1000, 1000, 2000).astype(float))
    mine = np.array([ol.lib().fso_hypot(float(a), float(b)) for a, b in zip(x, y)])
    ref = np.hypot(x, y)
    fin = np.isfinite(ref) & (ref > 1e-300)
    assert np.all(np.abs(mine[fin] - ref[fin]) <= 2 * np.spacing(ref[fin]))
    assert ol.lib().fso_hypot(0., 0.) == 0. and ol.lib().fso_hypot(3., 4.) == 5.


def _declared(header):
    txt = open(os.path.join(REPO, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", txt)))


def test_c_abi_exports_every_declared_symbol():
    """ the C-ABI libraries load on a CPU-only box and export exactly what
    include/*.h declares (no compute call is made) """
    for strict in (False, True):
        lib = _native.load_cuda_lib(strict)
        for sym in _declared("fsb200.h"):
            assert hasattr(lib, sym), sym
    assert sorted(_declared("fsb200.h")) == sorted(_native.CUDA_SYMBOLS)
    olib = _native.load_orbit_lib()
    for sym in _declared("fsb200_orbit.h"):
        assert hasattr(olib, sym), sym
    assert b"sm_100a" in _native.load_cuda_lib(False).fsb_build_info()
    assert b"fmad=off" in _native.load_cuda_lib(True).fsb_build_info()


def test_struct_layouts_match_headers():
    """ ctypes mirrors have the size the C compiler gives the structs """
    src = r'''
    #include "fsb200.h"
    #include "fsb200_orbit.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(fsb_stats), sizeof(fsb_std_desc),
                       sizeof(fsb_frame_desc), sizeof(fsb_orbit_xr), sizeof(fsb_postproc_desc),
                       sizeof(fsb_postproc_ext), offsetof(fsb_postproc_desc, df_k),
                       offsetof(fsb_postproc_ext, proj_hmoy)); return 0; }'''
    d = tempfile.mkdtemp()
    open(os.path.join(d, "s.c"), "w").write(src)
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(REPO, "include"),
                           os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
    sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    from fractalshades_b200 import postproc as fpp
    assert sizes == [ctypes.sizeof(_native.FsbStats), ctypes.sizeof(_native.FsbStdDesc),
                     ctypes.sizeof(_native.FsbFrameDesc), ctypes.sizeof(_native.OrbitXr),
                     ctypes.sizeof(fpp.FsbPostprocDesc), ctypes.sizeof(fpp.FsbPostprocExt),
                     fpp.FsbPostprocDesc.df_k.offset, fpp.FsbPostprocExt.proj_hmoy.offset]


def test_product_fails_loudly_without_gpu():
    """ no CPU fallback: on a box without a CUDA device compute calls raise """
    lib = _native.load_cuda_lib(False)
    if lib.fsb_device_count() > 0:
        pytest.skip("a GPU is present")
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=-1., y=0., dx=5., nx=64, xy_ratio=1., theta_deg=0.)
    f.calc_std_div(calc_name="c", subset=None, max_iter=100, M_divergence=1000.,
                   epsilon_stationnary=1e-3)
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        f.calc_raw("c")
    assert lib.fsb_frame_run(None, 10, None, None, None, None, None, None, None) < 0


def test_unsupported_features_are_refused():
    f = fsm.Mandelbrot(tempfile.mkdtemp())

    class Swirl(fsb.projection.Projection):      # a user mapping: no parametric form
        def __init__(self):
            pass
    with pytest.raises(NotImplementedError):
        f.zoom(x=0., y=0., dx=1., nx=64, xy_ratio=1., theta_deg=0., projection=Swirl())
    with pytest.raises(NotImplementedError):
        fsb.projection.Generic_mapping(lambda z: z, lambda z: 1.)
    with pytest.raises(ValueError):
        fsb.projection.Expmap(hmin=2., hmax=1.)
    with pytest.raises(ValueError):
        fsm.Burning_ship(tempfile.mkdtemp(), flavor="nope")
    with pytest.raises(TypeError):
        f.zoom(0., 0.)          # keyword-only, like the reference decorators


def test_native_orbit_known_answers():
    """ known answers of the reference's tests/test_FP_loop.py: the burning-
    ship orbit against 10 python iterations (:278-309) and the shallow
    Mandelbrot orbit against mpmath """
    import mpmath
    lib = _native.load_orbit_lib()
    n = 40
    orb = np.zeros(2 * (n + 1))
    cnt = ctypes.c_int64(0)
    buf = (_native.OrbitXr * 8)()
    x, y = "-1.7492046334590113", "-0.0002868466023466045"
    i = lib.fsb_orbit_mandelbrot(orb.ctypes.data, n, 2, 0, 2000., x.encode(), y.encode(),
                                 200, buf, 8, ctypes.byref(cnt))
    assert i == n + 1
    mpmath.mp.prec = 200
    c = mpmath.mpc(x, y)
    z = mpmath.mpc(0)
    for k in range(1, n + 1):
        z = z * z + c
        assert orb[2 * k] == float(z.real) and orb[2 * k + 1] == float(z.imag)
    a, b = mpmath.mpf("-1.75"), mpmath.mpf("-0.03")
    i = lib.fsb_orbit_burning_ship(orb.ctypes.data, 10, 1, 0, 2000., b"-1.75", b"-0.03",
                                   200, buf, 8, ctypes.byref(cnt))
    xx, yy = mpmath.mpf(0), mpmath.mpf(0)
    for k in range(1, min(i, 10) + 1):
        xx, yy = xx * xx - yy * yy + a, 2 * abs(xx * yy) - b
        assert orb[2 * k] == float(xx) and orb[2 * k + 1] == float(yy)
    # escape index
    i = lib.fsb_orbit_mandelbrot(orb.ctypes.data, n, 2, 0, 2000., b"1.0", b"1.0", 100,
                                 buf, 8, ctypes.byref(cnt))
    assert 0 < i <= 6


def test_fingerprint_file_is_written_atomically(tmp_path):
    """ several ranks may share one directory: a reader must never see a
    half-written fingerprint (the 8-GPU movie run raced on it) """
    import threading
    f = fsm.Mandelbrot(str(tmp_path))
    f.zoom(x=-1., y=0., dx=5., nx=64, xy_ratio=1.0, theta_deg=0.)
    fp = {"k": list(range(20000)), "name": "c"}
    f.save_fingerprint("c", fp)
    stop, errors = [False], []

    def reader():
        while not stop[0]:
            try:
                assert f.reload_fingerprint("c")["name"] == "c"
            except Exception as e:      # EOFError / UnpicklingError on a torn file
                errors.append(repr(e))
                return
    th = threading.Thread(target=reader)
    th.start()
    for _ in range(300):
        f.save_fingerprint("c", fp)
    stop[0] = True
    th.join()
    assert not errors, errors[:1]
    assert [p for p in os.listdir(os.path.dirname(f.fingerprint_path("c")))
            if p.endswith(".tmp")] == []


def test_pixel_grid_with_jitter_and_supersampling_matches_reference():
    """ chunk_pixel_pos (core.py:1767-1830) incl. the per-tile re-seeded jitter
    (default_rng(0)) and the supersampled grid: bit-identical to the live
    reference (fixture: tools/gen_golden_grid.py) """
    import hashlib
    import json
    g = np.load(os.path.join(REPO, "tests", "golden", "pixel_grid.npz"))
    for k, m in enumerate(json.loads(str(g["meta"]))):
        f = fsm.Mandelbrot(tempfile.mkdtemp())
        f.zoom(x=-0.5, y=0.1, dx=2.5, nx=m["nx"], xy_ratio=m["xy_ratio"], theta_deg=0.)
        shas, samples = [], []
        for cs in f.chunk_slices():
            p = np.ascontiguousarray(f.chunk_pixel_pos(cs, m["jitter"], m["supersampling"]))
            shas.append(hashlib.sha256(p.tobytes()).hexdigest())
            flat = p.ravel()
            samples.append(flat[np.linspace(0, flat.size - 1, 16).astype(np.int64)])
        assert np.array_equal(np.concatenate(samples), g[f"samples_{k}"]), m
        assert shas == m["sha"], m


def test_tile_axes_reproduce_the_pixel_grid():
    """ the per-tile axes sent to the grid calls (fsb_*_run_grid expands them on
    the device) give back chunk_pixel_pos bit for bit, ragged tiles included """
    import fractalshades_b200.models as fsm
    from fractalshades_b200.core import TileAxes
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=-0.5, y=0.1, dx=3., nx=450, xy_ratio=16 / 9., theta_deg=10.)
    tiles = list(f.chunk_slices())
    ta = TileAxes(f, tiles)
    assert ta.npts == f.nx * f.ny and ta.axes.dtype == np.float64
    off = 0
    for cs, (w, h) in zip(tiles, ta.shapes):
        x, y = ta.axes[off:off + w], ta.axes[off + w:off + w + h]
        off += w + h
        pos = f.chunk_pixel_pos(cs, False, None)
        assert pos.shape == (h, w)
        assert np.array_equal(pos.real, np.broadcast_to(x[None, :], (h, w)))
        assert np.array_equal(pos.imag, np.broadcast_to(y[:, None], (h, w)))
    assert off == ta.axes.shape[0]


def test_seam_validates_caller_buffers():
    """ the library copies nz * npts elements into Z: a buffer of another shape
    or type is refused before the call """
    import pytest
    from fractalshades_b200.core import check_outputs
    n = 100
    c = np.zeros(n, np.complex128)
    Z = np.zeros((2, n), np.complex128)
    U = np.zeros((1, n), np.int32)
    sr = np.zeros((1, n), np.int8)
    si = np.zeros((1, n), np.int32)
    check_outputs(n, Z, U, sr, si, 2, c)
    for bad in (dict(Z=np.zeros((1, n), np.complex128)), dict(Z=np.zeros((2, n - 1), np.complex128)),
                dict(Z=np.zeros((2, n), np.float32)), dict(U=np.zeros((1, n), np.int64)),
                dict(sr=np.zeros((1, n), np.int32)), dict(si=np.zeros((1, n + 1), np.int32)),
                dict(c=np.zeros(n, np.float64)), dict(Z=np.zeros((n, 2), np.complex128).T)):
        kw = dict(Z=Z, U=U, sr=sr, si=si, c=c)
        kw.update(bad)
        with pytest.raises(ValueError):
            check_outputs(n, kw["Z"], kw["U"], kw["sr"], kw["si"], 2, kw["c"])


def test_fingerprint_of_objects_is_stable():
    """ objects in a fingerprint (projection, ...) are described by class and parameters,
    not by an address: the stored results of a previous run are found again """
    from fractalshades_b200.core import _picklable
    from fractalshades_b200 import projection as prj
    a = _picklable({"zoom_kwargs": {"projection": prj.Expmap(0., 12.5, rotates_df=False), "nx": 64}})
    b = _picklable({"zoom_kwargs": {"projection": prj.Expmap(0., 12.5, rotates_df=False), "nx": 64}})
    c = _picklable({"zoom_kwargs": {"projection": prj.Expmap(0., 13.5, rotates_df=False), "nx": 64}})
    assert a == b and a != c
    assert " at 0x" not in a["zoom_kwargs"]["projection"]
    assert _picklable({"p": prj.Cartesian()}) == _picklable({"p": prj.Cartesian()})
    assert _picklable({"p": prj.Cartesian()}) != _picklable({"p": prj.Cartesian(expmap_seam=1.0)})
