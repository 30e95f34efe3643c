# -*- coding: utf-8 -*-
"""
Shared helpers of the parity tests: build a case's inputs with the PRODUCT's
host code (fractalshades_b200: orbit, Xrange scalars, pixel grid), then run it
through the CPU oracle and/or the CUDA path.
"""
import hashlib
import json
import os
import tempfile

import numpy as np

import oracle_lib as ol
from cases import CASES
import fractalshades_b200 as fsb
import fractalshades_b200.models as fsm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CLS = {"std_M2": fsm.Mandelbrot, "std_BS": fsm.Burning_ship,
        "perturb_M2": fsm.Perturbation_mandelbrot,
        "perturb_BS": fsm.Perturbation_burning_ship}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(name, mode):
    g = np.load(os.path.join(GOLDEN, f"{name}.{mode}.npz"))
    meta = json.loads(str(g["meta"]))
    return g, meta


def make_fractal(name, workdir=None):
    """ zoom() + option binding, WITHOUT touching the GPU """
    case = CASES[name]
    workdir = workdir or tempfile.mkdtemp(prefix="fsb_")
    cls = _CLS[case["kind"]]
    if "exponent" in case.get("init", {}):
        cls = (fsm.Perturbation_mandelbrot_N if case["kind"].startswith("perturb")
               else fsm.Mandelbrot_N)
    f = cls(workdir, **case.get("init", {}))
    from fractalshades_b200 import projection as _proj
    from cases import make_projection
    zoom = dict(x=case["x"], y=case["y"], dx=case["dx"], nx=case["nx"],
                xy_ratio=case.get("xy_ratio", 1.0),
                theta_deg=case.get("theta_deg", 0.),
                projection=make_projection(_proj, case.get("proj")),
                **case.get("skew", {}))
    if case["kind"].startswith("perturb"):
        zoom["precision"] = case["precision"]
    f.zoom(**zoom)
    # cases flagged newton=True use the reference's default flow (nucleus
    # search -> periodic reference); the flag is read when the orbit is built
    f._case_newton = bool(case.get("newton", False))
    step = (case.get("proj") or {}).get("step")
    if step is not None:      # one step of a stepped exponential zoom
        f.projection.set_exp_zoom_step(*step)
    return f, case


def bind_calc(f, case):
    """ What the calc_options decorator + calc_hook do, minus the device frame:
    returns the KernelSpec. """
    kw = dict(calc_name="c", subset=None, **case["calc"])
    ret = type(f).calc_std_div.__wrapped__(f, **kw)
    for k, v in kw.items():
        setattr(f, k, v)
    ret["set_state"]()(f)
    spec = ret["iterate"]()
    f._kernel_options = vars(spec).copy()
    return spec


def all_c_pix(f):
    return np.ascontiguousarray(np.concatenate(
        [np.ravel(f.chunk_pixel_pos(cs, False, None)) for cs in f.chunk_slices()]))


def host_tables(name):
    """ Per-frame tables from the product's host code (no GPU). """
    from fractalshades_b200 import settings
    f, case = make_fractal(name)
    bind_calc(f, case)
    old = settings.no_newton
    settings.no_newton = not f._case_newton
    try:
        t = f.frame_tables()
    finally:
        settings.no_newton = old
    return f, case, t


def oracle_fill_tables(t):
    """ dZndc / dZndz / BLA tables by the ORACLE builders (in place) """
    xr = t["xr_detect"]
    if t["kind"] == "perturb_M2":
        if t["calc_dzndc"]:
            t["dZndc"], t["dZndc_e"] = ol.dzndc_path_m2(
                t["Zn_path"], t["ref_index_xr"], t["ref_xr"], t["ref_xr_e"],
                t["ref_div_iter"], t["ref_order"], t.get("scale_deriv", t["dx"]),
                t.get("scale_deriv_e", t["dx_e"]), xr, t.get("nexp", 0))
        if t["calc_dzndz"]:
            t["dZndz"], t["dZndz_e"] = ol.dzndz_path_m2(
                t["Zn_path"], t["ref_index_xr"], t["ref_xr"], t["ref_xr_e"],
                t["ref_div_iter"], t["ref_order"], xr, t.get("nexp", 0))
        if t["bla_activated"]:
            (t["M_bla"], t["r_bla"], t["bla_len"], t["stages_bla"]
             ) = ol.make_bla_m2(t["Zn_path"], t["kc"], t["kc_e"], t["BLA_eps"],
                                t.get("nexp", 0))
    else:
        if t["calc_hessian"]:
            d4, e4 = ol.dzndc_path_bs(
                t["flavor"], t["Zn_path"], t["ref_index_xr"], t["refx_xr"],
                t["refx_xr_e"], t["refy_xr"], t["refy_xr_e"], t["ref_div_iter"],
                t["ref_order"], t.get("scale_deriv", t["dx"]),
                t.get("scale_deriv_e", t["dx_e"]), xr)
            for j, k in enumerate(("dXnda", "dXndb", "dYnda", "dYndb")):
                t[k] = d4[j]
                t[k + "_e"] = e4[j] if e4 is not None else None
        if t["bla_activated"]:
            (t["M_bla"], t["r_bla"], t["bla_len"], t["stages_bla"]
             ) = ol.make_bla_bs(t["flavor"], t["Zn_path"], t["kc"], t["kc_e"],
                                t["BLA_eps"])
    return t


def run_oracle(name, nthreads=0, det=False):
    """ (Z, U, stop_reason, stop_iter, extra) through the CPU oracle.  det:
    projections evaluated with the C library (False: what the reference runs,
    pinned by the fixtures) or with the platform-independent sequence shared
    with the CUDA library (True: the GPU parity target); see fs_oracle.h """
    case = CASES[name]
    if case["kind"].startswith("std"):
        f, case = make_fractal(name)
        c_pix0 = all_c_pix(f)
        pd = f.projection.c_abi_desc()
        c_pix = ol.project(dict(kind=pd.kind, hmoy=pd.hmoy, k_re=pd.pix_to_ht[0],
                                k_im=pd.pix_to_ht[1]), c_pix0, det)
        center = complex(f.x, f.y)
        if case["kind"] == "std_M2" and "exponent" in case.get("init", {}):
            # det: product chain (the CUDA definition) instead of the C
            # library's polar-form power the reference runs
            Z, U, sr, si = ol.std_mn(f.exponent, c_pix, center, f.dx, f.lin_mat,
                                     use_cpow=not det, nthreads=nthreads, **case["calc"])
        elif case["kind"] == "std_M2":
            Z, U, sr, si = ol.std_m2(c_pix, center, f.dx, f.lin_mat,
                                     nthreads=nthreads, **case["calc"])
        else:
            Z, U, sr, si = ol.std_bs(fsm.get_flavor_int(f.flavor), c_pix, center,
                                     f.dx, f.lin_mat, nthreads=nthreads,
                                     **case["calc"])
        return Z, U, sr, si, {"c_pix": c_pix0, "fractal": f}
    f, case, t = host_tables(name)
    oracle_fill_tables(t)
    c_pix = all_c_pix(f)
    Z, U, sr, si, cnt = ol.perturb(t, c_pix, nthreads, det)
    return Z, U, sr, si, {"c_pix": c_pix, "tables": t, "fractal": f,
                          "counters": cnt}


def run_gpu_case(name, strict, use_oracle_tables=False, tables=None):
    """ Same case through the CUDA path (C ABI via the product's Python).
    Perturbation: device frame built from the host tables; the library
    computes dZndc / BLA itself unless use_oracle_tables. """
    from fractalshades_b200 import settings, _native
    from fractalshades_b200.perturbation import create_frame
    case = CASES[name]
    if case["kind"].startswith("std"):
        f, case = make_fractal(name)
        spec = bind_calc(f, case)
        settings.strict_ieee = strict
        try:
            indep = f.get_cycle_indep_args(spec, spec)
            c_pix = all_c_pix(f)
            n = c_pix.shape[0]
            n_Z = len(f.codes[0])
            Z = np.zeros((n_Z, n), f.complex_type)
            U = np.zeros((0, n), np.int32)
            sr = -np.ones((1, n), np.int8)
            si = np.zeros((1, n), np.int32)
            rc = f.numba_cycle_call((c_pix, Z, U, sr, si), indep)
            assert rc == 0
        finally:
            settings.strict_ieee = False
        return Z, U, sr, si, {"stats": fsb.Fractal._last_stats}
    if tables is None:
        f, case, t = host_tables(name)
        c_pix = all_c_pix(f)
        if use_oracle_tables:
            oracle_fill_tables(t)
    else:
        t, c_pix = tables
    frame = create_frame(t, strict=strict, use_tables=use_oracle_tables)
    try:
        n = c_pix.shape[0]
        m2 = t["kind"] == "perturb_M2"
        Z = np.zeros((frame.nz, n), np.complex128 if m2 else np.float64)
        U = np.zeros((1, n), np.int32)
        sr = -np.ones((1, n), np.int8)
        si = np.zeros((1, n), np.int32)
        rc = frame.run(c_pix, Z, U, sr, si)
        assert rc == 0
        extra = {"stats": frame.last_stats, "setup_ms": frame.setup_ms()}
        if t["bla_activated"]:
            extra["bla"] = frame.get_bla()
        if (m2 and t["calc_dzndc"]) or ((not m2) and t["calc_hessian"]):
            extra["dzndc"] = frame.get_dzndc()
        if m2 and t["calc_dzndz"]:
            extra["dzndz"] = frame.get_dzndz()
    finally:
        frame.close()
    return Z, U, sr, si, extra


def same_bits(a, b):
    """ bit-for-bit equality treating any-NaN == any-NaN and -0 == +0 """
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if np.iscomplexobj(a):
        return same_bits(a.real, b.real) and same_bits(a.imag, b.imag)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def frac_same(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if np.iscomplexobj(a):
        ok = (((a.real == b.real) | (np.isnan(a.real) & np.isnan(b.real)))
              & ((a.imag == b.imag) | (np.isnan(a.imag) & np.isnan(b.imag))))
    else:
        ok = (a == b) | (np.isnan(a) & np.isnan(b))
    return float(np.mean(ok)) if ok.size else 1.0


def continuous_iter(kind, Zarr, n, M):
    """ nu = n - log(log|z| / log M) / log 2, the reference's Continuous_iter_pp
    for d = 2 (postproc.py:352-406,1001-1009), in fp64 """
    mod = np.hypot(Zarr[0], Zarr[1]) if kind.endswith("BS") else np.abs(Zarr[0])
    with np.errstate(all="ignore"):
        return n - np.log(np.log(mod) / np.log(M)) / np.log(2.)


def nu_within(kind, M, Za, na, Zb, nb, mask, tol=1e-9):
    """ fraction of the masked (matching, escaped) pixels whose continuous
    iteration agrees within `tol` relative; None when there is no such pixel """
    if mask.sum() == 0:
        return None
    a = continuous_iter(kind, Za[:, mask], na[0, mask].astype(float), M)
    b = continuous_iter(kind, Zb[:, mask], nb[0, mask].astype(float), M)
    fin = np.isfinite(a) & np.isfinite(b)
    if fin.sum() == 0:
        return None
    rel = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1.)
    return float(np.mean(rel < tol))
