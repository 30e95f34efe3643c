# -*- coding: utf-8 -*-
"""
GPU tests of the tile runtime around the kernels: the reference-facing Python
API end to end (zoom -> calc_std_div -> calc_raw -> memmaps in the reference's
layout), resume, interruption, thread-per-tile calls through the seam, and
size-independent properties at BASELINE.json's full 4K size.
"""
import concurrent.futures
import os
import tempfile
import threading
import time

import numpy as np
import pytest

import oracle_lib as ol
import parity_common as pc
import fractalshades_b200 as fsb
import fractalshades_b200.models as fsm
from fractalshades_b200 import settings, _native

pytestmark = pytest.mark.gpu


def _perturb_fractal(nx=300, directory=None, max_iter=50000):
    f = fsm.Perturbation_mandelbrot(directory or tempfile.mkdtemp())
    f.zoom(precision=30, x="-1.74928893611435556407228", y="0.", dx="5.e-20", nx=nx,
           xy_ratio=1.5, theta_deg=0.)
    f.calc_std_div(calc_name="c", subset=None, max_iter=max_iter, M_divergence=1e3,
                   epsilon_stationnary=1e-3, BLA_eps=1e-6, interior_detect=False,
                   calc_dzndc=True)
    return f


def test_calc_raw_memmaps_match_oracle_per_tile():
    """ public API end to end; the .arr files have the reference layout: tile
    slabs in chunk-rank order, row-major inside a tile (core.py:2362-2472) """
    settings.strict_ieee = True
    try:
        f = _perturb_fractal(nx=450)          # 3 x 2 tiles, ragged edges
        f.calc_raw("c")
    finally:
        settings.strict_ieee = False
    Zm = f.get_data_memmap("c", "Z", mode="r")
    Um = f.get_data_memmap("c", "U", mode="r")
    sim = f.get_data_memmap("c", "stop_iter", mode="r")
    srm = f.get_data_memmap("c", "stop_reason", mode="r")
    rep = f.get_report_memmap("c", mode="r")
    assert Zm.shape == (2, 450 * 300) and Zm.dtype == np.complex128
    assert Um.dtype == np.int32 and srm.dtype == np.int8 and sim.dtype == np.int32
    assert np.all(rep[:, 3] == 1)
    t = pc.oracle_fill_tables(dict(f._frame_tables))
    for rank, cs in enumerate(f.chunk_slices()):
        beg, end = int(rep[rank, 0]), int(rep[rank, 1])
        c_pix = np.ascontiguousarray(np.ravel(f.chunk_pixel_pos(cs, False, None)))
        Zo, Uo, sro, sio, _ = ol.perturb(t, c_pix)
        assert np.array_equal(sim[:, beg:end], sio)
        assert np.array_equal(srm[:, beg:end], sro)
        assert np.array_equal(Um[:, beg:end], Uo)
        assert pc.same_bits(np.array(Zm[:, beg:end]), Zo)
    # reload_data gives the tile back
    cs = f.chunk_from_rank(3)
    sub, c_pix, Z, U, sr, si = f.reload_data(cs, "c")
    assert Z.shape[1] == (cs[1] - cs[0]) * (cs[3] - cs[2])


def test_resume_skips_finished_tiles_and_tile_validator():
    f = _perturb_fractal(nx=450)
    f.calc_raw("c", tile_validator=lambda cs: f.chunk_rank(cs) % 2 == 0)
    rep = np.array(f.get_report_memmap("c", mode="r"))
    assert list(rep[:, 3]) == [1, 0, 1, 0, 1, 0]
    before = np.array(f.get_data_memmap("c", "stop_iter", mode="r"))
    f.calc_raw("c")
    rep = np.array(f.get_report_memmap("c", mode="r"))
    assert np.all(rep[:, 3] == 1)
    after = np.array(f.get_data_memmap("c", "stop_iter", mode="r"))
    done = np.zeros(after.shape[1], bool)
    for r in (0, 2, 4):
        done[rep[r, 0]:rep[r, 1]] = True
    assert np.array_equal(before[:, done], after[:, done])
    assert (after[:, ~done] > 0).all() and (before[:, ~done] == 0).all()
    # a second instance on the same directory finds the results (fingerprint)
    f2 = _perturb_fractal(nx=450, directory=f.directory)
    assert f2.res_available("c") and not f2._calc_data["c"]["need_new_mmap"]


def test_thread_per_tile_calls_through_the_seam():
    """ the reference calls numba_cycle_call once per tile from a thread pool
    (mthreading.py:48-68): the C ABI must be thread-safe """
    f = _perturb_fractal(nx=450)
    indep = f._calc_data["c"]["cycle_indep_args"]
    tiles = list(f.chunk_slices())

    def one(cs):
        dep, _ = f.get_cycling_dep_args("c", cs)
        rc = f.numba_cycle_call(dep, indep)
        assert rc == 0
        return dep
    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        res = list(ex.map(one, tiles * 3))
    f.calc_raw("c")
    rep = f.get_report_memmap("c", mode="r")
    sim = f.get_data_memmap("c", "stop_iter", mode="r")
    Zm = f.get_data_memmap("c", "Z", mode="r")
    for k, dep in enumerate(res):
        rank = k % len(tiles)
        beg, end = int(rep[rank, 0]), int(rep[rank, 1])
        assert np.array_equal(dep[4], sim[:, beg:end])
        assert pc.same_bits(dep[1], np.array(Zm[:, beg:end]))


def test_user_interruption_returns_code_1():
    """ core.py:2011-2020 / 2960-2961: a raised flag stops the computation and
    the call returns USER_INTERRUPTED """
    f = fsm.Perturbation_mandelbrot(tempfile.mkdtemp())
    v = fsb.VIEWS["deep_julia_2608"]
    f.zoom(precision=420, x=v["x"][:440], y=v["y"][:440], dx="1e-400", nx=2000,
           xy_ratio=1.0, theta_deg=0.)
    f.calc_std_div(calc_name="c", subset=None, max_iter=250000, M_divergence=1e3,
                   epsilon_stationnary=1e-3, BLA_eps=None, interior_detect=False,
                   calc_dzndc=True)            # no BLA: a long-running kernel
    indep = f._calc_data["c"]["cycle_indep_args"]
    c_pix = pc.all_c_pix(f)
    n = c_pix.shape[0]
    Z = np.zeros((2, n), np.complex128)
    U = np.zeros((1, n), np.int32)
    sr = -np.ones((1, n), np.int8)
    si = np.zeros((1, n), np.int32)
    timer = threading.Timer(0.02, f.raise_interruption)
    timer.start()
    t0 = time.time()
    rc = f.numba_cycle_call((c_pix, Z, U, sr, si), indep)
    dt = time.time() - t0
    assert rc == fsb.USER_INTERRUPTED
    assert (sr == -1).any()                    # unfinished pixels keep -1
    assert dt < 20.
    f.lower_interruption()
    assert f.numba_cycle_call((c_pix[:64], Z[:, :64].copy(), U[:, :64].copy(),
                               sr[:, :64].copy(), si[:, :64].copy()), indep) == 0


def test_standard_api_final_render_tile():
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=-0.75, y=0.1, dx=0.5, nx=300, xy_ratio=1.0, theta_deg=15.)
    f.calc_std_div(calc_name="s", subset=None, max_iter=2000, M_divergence=1000.,
                   epsilon_stationnary=1e-3)
    cs = (200, 300, 0, 200)
    ret = f.evaluate_rawdata_final("s", cs, {"jitter": 0.5, "supersampling": "2x2"})
    sub, c_pix, Z, U, sr, si = ret
    assert c_pix.shape[0] == 100 * 200 * 4 and Z.shape == (3, c_pix.shape[0])
    assert set(np.unique(sr)) <= {0, 1, 2}


@pytest.fixture(scope="module")
def full_4k():
    """ BASELINE config 2 at its full 4K size, default build """
    import bench
    w = bench.WORKLOADS["config2"]
    f = bench.make_fractal(w)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    indep = f._calc_data["bench"]["cycle_indep_args"]
    c_pix = bench.frame_c_pix(f)
    n = c_pix.shape[0]
    out = []
    for _ in range(2):
        Z = np.zeros((2, n), np.complex128)
        U = np.zeros((1, n), np.int32)
        sr = -np.ones((1, n), np.int8)
        si = np.zeros((1, n), np.int32)
        assert f.numba_cycle_call((c_pix, Z, U, sr, si), indep) == 0
        out.append((Z, U, sr, si))
    return f, indep, c_pix, out


def test_full_size_properties_4k(full_4k):
    """ size-independent properties at 3840x2160: determinism (two runs are
    identical), every pixel finished, counters consistent, a permutation of
    the point list permutes the outputs, and three whole tiles equal the
    oracle within the default-build tolerance """
    f, indep, c_pix, out = full_4k
    (Z, U, sr, si), (Z2, U2, sr2, si2) = out
    n = c_pix.shape[0]
    assert n == 3840 * 2160 and f.chunks_count == 220
    assert np.array_equal(si, si2) and np.array_equal(sr, sr2)
    assert pc.same_bits(Z, Z2) and np.array_equal(U, U2)
    assert (sr >= 0).all() and (si > 0).all()
    st = fsb.Fractal._last_stats
    assert st["sum_stop_iter"] == int(si.sum(dtype=np.int64))
    assert st["n_iter_exec"] + st["n_bla_steps"] <= st["sum_stop_iter"]
    # permutation equivariance (work distribution does not leak into results)
    rg = np.random.default_rng(0)
    sel = rg.permutation(n)[:500000]
    cp = np.ascontiguousarray(c_pix[sel])
    Zp = np.zeros((2, sel.size), np.complex128)
    Up = np.zeros((1, sel.size), np.int32)
    srp = -np.ones((1, sel.size), np.int8)
    sip = np.zeros((1, sel.size), np.int32)
    assert f.numba_cycle_call((cp, Zp, Up, srp, sip), indep) == 0
    assert np.array_equal(sip, si[:, sel]) and pc.same_bits(Zp, Z[:, sel])
    # whole tiles against the oracle
    t = pc.oracle_fill_tables(dict(f._frame_tables))
    offs = np.cumsum([0] + [(c[1] - c[0]) * (c[3] - c[2]) for c in f.chunk_slices()])
    for rank in (0, 110, 219):
        beg, end = offs[rank], offs[rank + 1]
        Zo, Uo, sro, sio, _ = ol.perturb(t, np.ascontiguousarray(c_pix[beg:end]))
        same = (sio == si[:, beg:end])[0] & (sro == sr[:, beg:end])[0]
        assert same.mean() >= 0.999, (rank, same.mean())
        rel = np.abs(Z[0, beg:end][same] - Zo[0][same]) / np.abs(Zo[0][same])
        assert np.median(rel) < 1e-9


# ---------------------------------------------------------------------------
# tile-list calls (8 x 4 patch mapping) == flat calls, bit for bit
_TILE_SHAPES = [(200, 200), (1, 1), (7, 3), (9, 5), (40, 160), (8, 4), (33, 1), (1, 37),
                (16, 9)]


def _split_tiles(n, shapes):
    """ tile shapes covering exactly n points (the last one is a 1-row strip) """
    out, left = [], n
    k = 0
    while left > 0:
        w, h = shapes[k % len(shapes)]
        k += 1
        if w * h > left:
            out.append((left, 1))
            break
        out.append((w, h))
        left -= w * h
    return out


@pytest.mark.parametrize("name", ["p_M2_E20", "p_M2_deep1000_xr", "p_M2_int_E11", "p_BS_f1_E30_skew", "p_BS_f1_E500_xr"])
@pytest.mark.parametrize("strict", [True, False])
def test_tile_list_calls_equal_flat_calls(name, strict):
    from cases import CASES
    if name not in CASES:
        pytest.skip("case not defined")
    from fractalshades_b200.perturbation import create_frame
    f, case, t = pc.host_tables(name)
    c_pix = pc.all_c_pix(f)
    n = c_pix.shape[0]
    tiles = _split_tiles(n, _TILE_SHAPES)
    assert sum(w * h for w, h in tiles) == n
    frame = create_frame(t, strict=strict)
    try:
        m2 = t["kind"] == "perturb_M2"
        outs = []
        for tl in (None, tiles):
            Z = np.zeros((frame.nz, n), np.complex128 if m2 else np.float64)
            U = np.zeros((1, n), np.int32)
            sr = -np.ones((1, n), np.int8)
            si = np.zeros((1, n), np.int32)
            assert frame.run(c_pix, Z, U, sr, si, tiles=tl) == 0
            outs.append((Z, U, sr, si, dict(frame.last_stats)))
    finally:
        frame.close()
    a, b = outs
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[3], b[3])
    assert a[0].tobytes() == b[0].tobytes()
    for k in ("n_iter_exec", "n_bla_steps", "n_rebase", "sum_stop_iter"):
        assert a[4][k] == b[4][k]


def test_tile_list_standard_loop_and_bad_tiles():
    f = fsm.Mandelbrot(tempfile.mkdtemp())
    f.zoom(x=-1.0, y=0.0, dx=5.0, nx=230, xy_ratio=1.0, theta_deg=0.)
    f.calc_std_div(calc_name="s", subset=None, max_iter=2000, M_divergence=1000.,
                   epsilon_stationnary=0.001)
    indep = f._calc_data["s"]["cycle_indep_args"]
    shapes, pix = [], []
    for cs in f.chunk_slices():
        pos = f.chunk_pixel_pos(cs, False, None)
        shapes.append((pos.shape[1], pos.shape[0]))
        pix.append(np.ravel(pos))
    c_pix = np.ascontiguousarray(np.concatenate(pix))
    n = c_pix.shape[0]
    outs = []
    for tl in (None, shapes):
        Z = np.zeros((3, n), np.complex128)
        U = np.zeros((0, n), np.int32)
        sr = -np.ones((1, n), np.int8)
        si = np.zeros((1, n), np.int32)
        assert f.numba_cycle_call((c_pix, Z, U, sr, si), indep, tiles=tl) == 0
        outs.append((Z, sr, si))
    assert outs[0][0].tobytes() == outs[1][0].tobytes()
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    with pytest.raises(ValueError):
        f.numba_cycle_call((c_pix, Z, U, sr, si), indep, tiles=[(10, 10)])


def test_grid_calls_match_cpix_calls():
    """ fsb_*_run_grid (pixel grid expanded on the device from per-tile axes) =
    the c_pix calls, bit for bit, for a perturbation frame and a standard one """
    from fractalshades_b200.core import TileAxes
    for f in (_perturb_fractal(nx=450), None):
        if f is None:
            f = fsm.Mandelbrot(tempfile.mkdtemp())
            f.zoom(x=-0.7, y=0.2, dx=2.5, nx=450, xy_ratio=1.5, theta_deg=20.)
            f.calc_std_div(calc_name="c", subset=None, max_iter=2000, M_divergence=1e3,
                           epsilon_stationnary=1e-3)
        indep = f._calc_data["c"]["cycle_indep_args"]
        state = f._calc_data["c"]["state"]
        tiles = list(f.chunk_slices())
        ta = TileAxes(f, tiles)
        n = ta.npts
        n_Z, n_U = len(state.codes[0]), len(state.codes[1])
        out = []
        for src in ("grid", "c_pix"):
            Z = np.zeros((n_Z, n), state.complex_type)
            U = np.zeros((n_U, n), np.int32)
            sr = -np.ones((1, n), np.int8)
            si = np.zeros((1, n), np.int32)
            if src == "grid":
                rc = f.numba_cycle_call((ta, Z, U, sr, si), indep)
            else:
                c_pix = np.ascontiguousarray(np.concatenate(
                    [np.ravel(f.chunk_pixel_pos(cs, False, None)) for cs in tiles]))
                rc = f.numba_cycle_call((c_pix, Z, U, sr, si), indep, tiles=ta.shapes)
            assert rc == 0
            out.append((Z, U, sr, si))
        (Z0, U0, sr0, si0), (Z1, U1, sr1, si1) = out
        assert np.array_equal(si0, si1) and np.array_equal(sr0, sr1) and np.array_equal(U0, U1)
        assert pc.same_bits(Z0, Z1)
        assert (sr0 >= 0).all()
