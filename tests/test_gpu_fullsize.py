# -*- coding: utf-8 -*-
"""
Parity at BASELINE.json's FULL sizes: whole tiles of every
configured frame, through the C ABI, against the oracle run on the same pixels.

  config 1  Mandelbrot 800x800, max_iter 5000: the WHOLE frame, bit-exact
            (the standard loop is IEEE-strict in both builds)
  config 2  1e-250, 4K, max_iter 1e6            | 4 whole 200x200 tiles each:
  config 3  1e-1000 (Xrange), 4K, max_iter 1e7  |   -fmad=false build bit-exact
  config 4  burning ship 1e-500, hessian, 4K    |   (ints, U and Z), default build
  config 5  one frame of the 8K zoom movie      |   >= 99.9 % + nu within 1e-9

The measured default-build match rates are printed and written to
gpurun_out/parity_rates.json (committed copy: profiles/parity_rates_r2.json).
"""
import json
import os
import tempfile

import numpy as np
import pytest

import oracle_lib as ol
import parity_common as pc
import fractalshades_b200.models as fsm
from fractalshades_b200 import settings

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RATES = {}


def _dump_rates():
    out = os.path.join(REPO, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_rates.json"), "w") as fh:
        json.dump(RATES, fh, indent=1, sort_keys=True)


def _tiles_c_pix(f, ranks):
    tiles = [f.chunk_from_rank(r) for r in ranks]
    pix = [np.ravel(f.chunk_pixel_pos(cs, False, None)) for cs in tiles]
    shapes = [(cs[1] - cs[0], cs[3] - cs[2]) for cs in tiles]
    return np.ascontiguousarray(np.concatenate(pix)), shapes


def _run(f, calc, c_pix, shapes, strict):
    settings.strict_ieee = strict
    try:
        f.calc_std_div(calc_name=calc, subset=None, **f._bench_calc)
        indep = f._calc_data[calc]["cycle_indep_args"]
        state = f._calc_data[calc]["state"]
        n = c_pix.shape[0]
        Z = np.zeros((len(state.codes[0]), n), state.complex_type)
        U = np.zeros((len(state.codes[1]), n), np.int32)
        sr = -np.ones((1, n), np.int8)
        si = np.zeros((1, n), np.int32)
        assert f.numba_cycle_call((c_pix, Z, U, sr, si), indep, tiles=shapes) == 0
        f._release_indep_args(indep)
    finally:
        settings.strict_ieee = False
    return Z, U, sr, si


def _check_perturb(name, f, kind, ranks, M):
    """ strict build == oracle bit for bit; default build within the north-star
    tolerance; rates recorded """
    c_pix, shapes = _tiles_c_pix(f, ranks)
    Zs, Us, srs, sis = _run(f, "s", c_pix, shapes, strict=True)
    t = pc.oracle_fill_tables(dict(f._frame_tables))
    Zo, Uo, sro, sio, cnt = ol.perturb(t, c_pix)
    assert (srs >= 0).all()
    assert np.array_equal(sis, sio) and np.array_equal(srs, sro) and np.array_equal(Us, Uo)
    assert pc.same_bits(Zs, Zo)
    Zd, Ud, srd, sid = _run(f, "d", c_pix, shapes, strict=False)
    same = (sid == sio)[0] & (srd == sro)[0]
    esc = same & (sro[0] == 1)
    nu = pc.nu_within(kind, M, Zd, sid, Zo, sio, esc)
    RATES[name] = {
        "tiles": list(map(int, ranks)), "points": int(c_pix.shape[0]),
        "strict_build_bit_exact": True,
        "default_build_stop_iter_and_reason_equal": float(same.mean()),
        "default_build_nu_within_1e-9": nu,
        "max_stop_iter": int(sio.max()), "escaped_fraction": float((sro[0] == 1).mean()),
        "oracle_exec_iterations": int(cnt[0]), "oracle_bla_steps": int(cnt[1]),
    }
    _dump_rates()
    print("\nPARITY", name, json.dumps(RATES[name]))
    assert same.mean() >= 0.999, (name, same.mean())
    assert nu is None or nu >= 0.995, (name, nu)


def _bench_fractal(wname, nx=None):
    import bench
    w = bench.WORKLOADS[wname]
    f = bench.make_fractal(w, nx)
    f._bench_calc = w["calc"]
    return f, w


def test_config1_whole_frame_bit_exact():
    f, w = _bench_fractal("config1")
    assert (f.nx, f.ny) == (800, 800)
    ranks = list(range(f.chunks_count))
    c_pix, shapes = _tiles_c_pix(f, ranks)
    f._frame_tables = None
    for strict in (True, False):
        Z, U, sr, si = _run(f, "c", c_pix, shapes, strict)
        Zo, Uo, sro, sio = ol.std_m2(c_pix, complex(f.x, f.y), float(f.dx), f.lin_mat, **w["calc"])
        assert np.array_equal(si, sio) and np.array_equal(sr, sro)
        assert pc.same_bits(Z, Zo)
    RATES["config1"] = {"points": int(c_pix.shape[0]), "both_builds_bit_exact": True,
                        "max_stop_iter": int(sio.max())}
    _dump_rates()
    print("\nPARITY config1", json.dumps(RATES["config1"]))


@pytest.mark.parametrize("wname,kind", [("config2", "perturb_M2"), ("config3", "perturb_M2"),
                                        ("config4", "perturb_BS")])
def test_full_size_tiles(wname, kind):
    f, w = _bench_fractal(wname)
    assert f.nx == 3840
    n = f.chunks_count
    ranks = sorted({0, n // 3, n // 2 + 3, n - 1})
    _check_perturb(wname, f, kind, ranks, float(w["calc"]["M_divergence"]))


def test_config5_movie_frame_tile():
    """ frame 20 of the 64-frame 8K movie (dx = 1e-641.7, Xrange): two tiles """
    import mpmath
    from fractalshades_b200 import movie
    from fractalshades_b200.views import VIEWS
    settings.no_newton = True
    v = VIEWS["deep_julia_2608"]
    digits = 700
    seq = movie.ZoomSequence(
        fsm.Perturbation_mandelbrot, tempfile.mkdtemp(), x=v["x"][:digits + 20],
        y=v["y"][:digits + 20], dx_start="1e-10", dx_end="1e-2000", n_frames=64, nx=7680,
        xy_ratio=16 / 9., precision=digits,
        calc_kwargs=dict(max_iter=3000000, M_divergence=1e3, epsilon_stationnary=1e-3,
                         BLA_eps=1e-6, interior_detect=False, calc_dzndc=True))
    f = seq._fractal(20)
    assert (f.nx, f.ny) == (7680, 4320)
    assert float(mpmath.log10(seq.widths[20])) < -600
    f._bench_calc = seq.calc_kwargs
    n = f.chunks_count
    _check_perturb("config5_frame20", f, "perturb_M2", [n // 2 + 7, n - 1], 1e3)
