# -*- coding: utf-8 -*-
"""
ctypes binding of oracle/liboracle.so (CPU oracle, oracle/fs_oracle.h).

TEST INFRASTRUCTURE: imported by tests/, tools/gen_golden.py,
__graft_entry__.smoke() and the cpu_baseline legs of bench.py only.

A perturbation frame is exchanged as a plain dict of numpy arrays / scalars
("tables dict"), see `tables_from_reference` (tools/gen_golden.py) and
`fractalshades_b200.perturbation.PerturbationFractal.frame_tables`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p


class FrameM2(ctypes.Structure):
    _fields_ = [
        ("L", c_i64), ("Zn_path", c_vp), ("dZndc", c_vp), ("dZndc_e", c_vp),
        ("dZndz", c_vp), ("dZndz_e", c_vp), ("n_xr", c_i64),
        ("ref_index_xr", c_vp), ("ref_xr", c_vp), ("ref_xr_e", c_vp),
        ("ref_div_iter", c_i64), ("ref_order", c_i64), ("drift", c_dbl * 2),
        ("drift_e", c_i32), ("lin_scale_e", c_i32), ("lin_scale", c_dbl),
        ("lin_mat", c_dbl * 4), ("M_bla", c_vp), ("r_bla", c_vp),
        ("bla_len", c_i64), ("stages_bla", c_i32), ("xr_detect", c_i32),
        ("bla_activated", c_i32), ("calc_dzndc", c_i32), ("calc_dzndz", c_i32),
        ("calc_orbit", c_i32), ("backshift", c_i64), ("max_iter", c_i64),
        ("M_divergence_sq", c_dbl), ("epsilon_stationnary_sq", c_dbl),
        ("nexp", c_i32), ("use_cpow", c_i32),
    ]


class FrameBS(ctypes.Structure):
    _fields_ = [
        ("L", c_i64), ("Zn_path", c_vp),
        ("dXnda", c_vp), ("dXndb", c_vp), ("dYnda", c_vp), ("dYndb", c_vp),
        ("dXnda_e", c_vp), ("dXndb_e", c_vp), ("dYnda_e", c_vp), ("dYndb_e", c_vp),
        ("n_xr", c_i64), ("ref_index_xr", c_vp), ("refx_xr", c_vp),
        ("refy_xr", c_vp), ("refx_xr_e", c_vp), ("refy_xr_e", c_vp),
        ("ref_div_iter", c_i64), ("ref_order", c_i64),
        ("driftx", c_dbl), ("drifty", c_dbl), ("driftx_e", c_i32),
        ("drifty_e", c_i32), ("lin_scale", c_dbl), ("lin_scale_e", c_i32),
        ("flavor", c_i32), ("lin_mat", c_dbl * 4), ("M_bla", c_vp),
        ("r_bla", c_vp), ("bla_len", c_i64), ("stages_bla", c_i32),
        ("xr_detect", c_i32), ("bla_activated", c_i32), ("calc_hessian", c_i32),
        ("calc_orbit", c_i32), ("_pad", c_i32), ("backshift", c_i64),
        ("max_iter", c_i64), ("M_divergence_sq", c_dbl),
    ]


def build(force=False):
    """ Compile oracle/liboracle.so (g++, seconds). """
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("fs_oracle.cpp", "fs_oracle.h")]
    if (not force and os.path.exists(so)
            and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in src)):
        return so
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], env=env,
                          stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.fso_hypot.restype = c_dbl
        _LIB.fso_hypot.argtypes = [c_dbl, c_dbl]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data


def _c128(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.complex128)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------------------
# standard loops
def std_m2(c_pix, center, dx, lin_mat, max_iter, M_divergence,
           epsilon_stationnary, calc_d2zndc2=False, calc_orbit=False,
           backshift=0, nthreads=0):
    c_pix = _c128(c_pix)
    n = c_pix.shape[0]
    nz = 3 + int(calc_d2zndc2) + int(calc_orbit)
    Z = np.zeros((nz, n), np.complex128)
    sr = np.full((1, n), -1, np.int8)
    si = np.zeros((1, n), np.int32)
    lm = _f64(lin_mat).ravel()
    lib().fso_std_m2(c_i64(n), c_vp(_p(c_pix)), c_dbl(center.real),
                     c_dbl(center.imag), c_dbl(dx), c_vp(_p(lm)),
                     c_i64(max_iter), c_dbl(M_divergence ** 2),
                     c_dbl(epsilon_stationnary ** 2), int(calc_d2zndc2),
                     int(calc_orbit), c_i64(backshift), c_vp(_p(Z)),
                     c_vp(_p(sr)), c_vp(_p(si)), int(nthreads))
    return Z, np.zeros((0, n), np.int32), sr, si


def std_mn(nexp, c_pix, center, dx, lin_mat, max_iter, M_divergence,
           epsilon_stationnary, calc_d2zndc2=False, calc_orbit=False, backshift=0,
           use_cpow=True, nthreads=0):
    """ Mandelbrot_N; use_cpow: see fs_oracle.h """
    c_pix = _c128(c_pix)
    n = c_pix.shape[0]
    Z = np.zeros((3 + int(calc_d2zndc2) + int(calc_orbit), n), np.complex128)
    sr = np.full((1, n), -1, np.int8)
    si = np.zeros((1, n), np.int32)
    lm = _f64(lin_mat).ravel()
    lib().fso_std_mn(int(nexp), int(bool(use_cpow)), c_i64(n), c_vp(_p(c_pix)),
                     c_dbl(center.real), c_dbl(center.imag), c_dbl(dx), c_vp(_p(lm)),
                     c_i64(max_iter), c_dbl(M_divergence ** 2),
                     c_dbl(epsilon_stationnary ** 2), int(calc_d2zndc2), int(calc_orbit),
                     c_i64(backshift), c_vp(_p(Z)),
                     c_vp(_p(sr)), c_vp(_p(si)), int(nthreads))
    return Z, np.zeros((0, n), np.int32), sr, si


def std_bs(flavor, c_pix, center, dx, lin_mat, max_iter, M_divergence,
           calc_orbit=False, backshift=0, nthreads=0):
    c_pix = _c128(c_pix)
    n = c_pix.shape[0]
    nz = 6 + 2 * int(calc_orbit)
    Z = np.zeros((nz, n), np.float64)
    sr = np.full((1, n), -1, np.int8)
    si = np.zeros((1, n), np.int32)
    lm = _f64(lin_mat).ravel()
    lib().fso_std_bs(int(flavor), c_i64(n), c_vp(_p(c_pix)),
                     c_dbl(center.real), c_dbl(center.imag), c_dbl(dx),
                     c_vp(_p(lm)), c_i64(max_iter), c_dbl(M_divergence ** 2),
                     int(calc_orbit), c_i64(backshift), c_vp(_p(Z)),
                     c_vp(_p(sr)), c_vp(_p(si)), int(nthreads))
    return Z, np.zeros((0, n), np.int32), sr, si


# ---------------------------------------------------------------------------
# perturbation loops
def _frame_m2(t, keep):
    f = FrameM2()
    Zn = _c128(t["Zn_path"]); keep.append(Zn)
    f.L = Zn.shape[0]
    f.Zn_path = _p(Zn)
    for k, conv in (("dZndc", _c128), ("dZndc_e", _i32), ("dZndz", _c128),
                    ("dZndz_e", _i32), ("ref_index_xr", _i32),
                    ("ref_xr", _c128), ("ref_xr_e", _i32), ("M_bla", _c128),
                    ("r_bla", _f64)):
        a = conv(t.get(k)); keep.append(a)
        setattr(f, k, _p(a))
    f.n_xr = 0 if t.get("ref_index_xr") is None else len(t["ref_index_xr"])
    f.ref_div_iter = int(t["ref_div_iter"])
    f.ref_order = int(t["ref_order"])
    d = complex(t["drift"])
    f.drift[0], f.drift[1] = d.real, d.imag
    f.drift_e = int(t["drift_e"])
    f.lin_scale = float(t["lin_scale"])
    f.lin_scale_e = int(t["lin_scale_e"])
    lm = np.asarray(t["lin_mat"], np.float64).ravel()
    for i in range(4):
        f.lin_mat[i] = lm[i]
    f.bla_len = int(t.get("bla_len") or 0)
    f.stages_bla = int(t.get("stages_bla") or 0)
    for k in ("xr_detect", "bla_activated", "calc_dzndc", "calc_dzndz",
              "calc_orbit"):
        setattr(f, k, int(bool(t.get(k, False))))
    f.backshift = int(t.get("backshift", 0))
    f.max_iter = int(t["max_iter"])
    f.M_divergence_sq = float(t["M_divergence"]) ** 2
    f.epsilon_stationnary_sq = float(t.get("epsilon_stationnary", 0.)) ** 2
    f.nexp = int(t.get("nexp", 0) or 0)       # Perturbation_mandelbrot_N
    f.use_cpow = int(bool(t.get("use_cpow", False)))
    return f


def nz_m2(t):
    return (1 + int(bool(t.get("calc_dzndz"))) + int(bool(t.get("calc_dzndc")))
            + int(bool(t.get("calc_orbit"))))


def perturb_m2(t, c_pix, nthreads=0):
    keep = []
    f = _frame_m2(t, keep)
    c_pix = _c128(c_pix)
    n = c_pix.shape[0]
    Z = np.zeros((nz_m2(t), n), np.complex128)
    U = np.zeros((1, n), np.int32)
    sr = np.full((1, n), -1, np.int8)
    si = np.zeros((1, n), np.int32)
    cnt = np.zeros(3, np.int64)
    lib().fso_perturb_m2(ctypes.byref(f), c_i64(n), c_vp(_p(c_pix)),
                         c_vp(_p(Z)), c_vp(_p(U)), c_vp(_p(sr)), c_vp(_p(si)),
                         int(nthreads), c_vp(_p(cnt)))
    return Z, U, sr, si, cnt


def _frame_bs(t, keep):
    f = FrameBS()
    Zn = _c128(t["Zn_path"]); keep.append(Zn)
    f.L = Zn.shape[0]
    f.Zn_path = _p(Zn)
    for k in ("dXnda", "dXndb", "dYnda", "dYndb", "refx_xr", "refy_xr",
              "M_bla", "r_bla"):
        a = _f64(t.get(k)); keep.append(a)
        setattr(f, k, _p(a))
    for k in ("dXnda_e", "dXndb_e", "dYnda_e", "dYndb_e", "ref_index_xr",
              "refx_xr_e", "refy_xr_e"):
        a = _i32(t.get(k)); keep.append(a)
        setattr(f, k, _p(a))
    f.n_xr = 0 if t.get("ref_index_xr") is None else len(t["ref_index_xr"])
    f.ref_div_iter = int(t["ref_div_iter"])
    f.ref_order = int(t["ref_order"])
    f.driftx, f.driftx_e = float(t["driftx"]), int(t["driftx_e"])
    f.drifty, f.drifty_e = float(t["drifty"]), int(t["drifty_e"])
    f.lin_scale = float(t["lin_scale"])
    f.lin_scale_e = int(t["lin_scale_e"])
    f.flavor = int(t["flavor"])
    lm = np.asarray(t["lin_mat"], np.float64).ravel()
    for i in range(4):
        f.lin_mat[i] = lm[i]
    f.bla_len = int(t.get("bla_len") or 0)
    f.stages_bla = int(t.get("stages_bla") or 0)
    for k in ("xr_detect", "bla_activated", "calc_hessian", "calc_orbit"):
        setattr(f, k, int(bool(t.get(k, False))))
    f.backshift = int(t.get("backshift", 0))
    f.max_iter = int(t["max_iter"])
    f.M_divergence_sq = float(t["M_divergence"]) ** 2
    return f


def nz_bs(t):
    return 2 + 4 * int(bool(t.get("calc_hessian"))) + 2 * int(bool(t.get("calc_orbit")))


def perturb_bs(t, c_pix, nthreads=0):
    keep = []
    f = _frame_bs(t, keep)
    c_pix = _c128(c_pix)
    n = c_pix.shape[0]
    Z = np.zeros((nz_bs(t), n), np.float64)
    U = np.zeros((1, n), np.int32)
    sr = np.full((1, n), -1, np.int8)
    si = np.zeros((1, n), np.int32)
    cnt = np.zeros(3, np.int64)
    lib().fso_perturb_bs(ctypes.byref(f), c_i64(n), c_vp(_p(c_pix)),
                         c_vp(_p(Z)), c_vp(_p(U)), c_vp(_p(sr)), c_vp(_p(si)),
                         int(nthreads), c_vp(_p(cnt)))
    return Z, U, sr, si, cnt


def perturb(t, c_pix, nthreads=0, det=False):
    """ projection of the pixels, the loop, then the dz/dc modifier.  det: see
    fs_oracle.h (C-library or platform-independent exp / sin / cos) """
    pj = t.get("proj")
    pix = project(pj, c_pix, det)
    if t["kind"] == "perturb_M2":
        # power N with calc_orbit: zn ** N as the reference runs it (C library) or as the
        # CUDA library defines it (product chain), like the projections
        out = perturb_m2(dict(t, use_cpow=not det), pix, nthreads)
    else:
        out = perturb_bs(t, pix, nthreads)
    apply_modifier(t, out[0], modifier(pj, c_pix, det))
    return out


# ---------------------------------------------------------------------------
# projections (fso_proj_expmap & co); `pj` = the "proj" dict of the frame tables
def project(pj, c_pix, det):
    c_pix = _c128(c_pix)
    if pj is None or int(pj["kind"]) == 0:
        return c_pix
    out = np.empty_like(c_pix)
    lib().fso_proj_expmap(c_i64(c_pix.shape[0]), c_vp(_p(c_pix)), c_dbl(pj["hmoy"]),
                          c_dbl(pj["k_re"]), c_dbl(pj["k_im"]), int(bool(det)),
                          c_vp(_p(out)))
    return out


def modifier(pj, c_pix, det):
    """ proj_dzndc_modifier(c_pix) for every pixel, or None """
    if pj is None or int(pj["dzndc_modifier"]) == 0:
        return None
    c_pix = _c128(c_pix)
    out = np.empty(c_pix.shape[0], np.float64)
    if int(pj["dzndc_modifier"]) == 1:
        lib().fso_modifier_expmap(c_i64(out.size), c_vp(_p(c_pix)), c_dbl(pj["k_re"]),
                                  c_dbl(pj["k_im"]), c_dbl(pj["mod_param"]),
                                  int(bool(det)), c_vp(_p(out)))
    else:
        lib().fso_modifier_seam(c_i64(out.size), c_vp(_p(c_pix)), c_dbl(pj["mod_param"]),
                                int(bool(det)), c_vp(_p(out)))
    return out


def apply_modifier(t, Z, mod):
    """ perturbation.py:1387-1388 / 1772-1776 on the derivative rows of Z """
    if mod is None:
        return
    n = Z.shape[1]
    if t["kind"] == "perturb_M2":
        if t.get("calc_dzndc"):
            row = 1 + int(bool(t.get("calc_dzndz")))
            r = np.ascontiguousarray(Z[row])
            lib().fso_apply_modifier_c(c_i64(n), c_vp(_p(r)), c_vp(_p(mod)))
            Z[row] = r
    elif t.get("calc_hessian"):
        for row in range(2, 6):
            r = np.ascontiguousarray(Z[row])
            lib().fso_apply_modifier_f(c_i64(n), c_vp(_p(r)), c_vp(_p(mod)))
            Z[row] = r


def det_exp(x):
    lib().fso_det_exp.restype = c_dbl
    return np.array([lib().fso_det_exp(c_dbl(v)) for v in np.ravel(x)])


def det_sincos(x):
    s, c = c_dbl(), c_dbl()
    out = np.empty((np.size(x), 2))
    for i, v in enumerate(np.ravel(x)):
        lib().fso_det_sincos(c_dbl(v), ctypes.byref(s), ctypes.byref(c))
        out[i] = s.value, c.value
    return out


# ---------------------------------------------------------------------------
# per-frame tables
def make_bla_m2(Zn_path, kc, kc_e, eps, nexp=0):
    Zn = _c128(Zn_path)
    L = Zn.shape[0]
    bla_len = 2 * (L // 8)
    M = np.zeros(2 * bla_len, np.complex128)
    r = np.zeros(bla_len, np.float64)
    stages = lib().fso_make_bla_mn(int(nexp or 0), c_vp(_p(Zn)), c_i64(L), c_dbl(kc),
                                   c_i32(kc_e), c_dbl(eps), c_vp(_p(M)),
                                   c_vp(_p(r)))
    return M, r, bla_len, stages


def make_bla_bs(flavor, Zn_path, kc, kc_e, eps):
    Zn = _c128(Zn_path)
    L = Zn.shape[0]
    bla_len = 2 * (L // 8)
    M = np.zeros(8 * bla_len, np.float64)
    r = np.zeros(bla_len, np.float64)
    stages = lib().fso_make_bla_bs(int(flavor), c_vp(_p(Zn)), c_i64(L),
                                   c_dbl(kc), c_i32(kc_e), c_dbl(eps),
                                   c_vp(_p(M)), c_vp(_p(r)))
    return M, r, bla_len, stages


def dzndc_path_m2(Zn_path, ref_index_xr, ref_xr, ref_xr_e, ref_div_iter,
                  ref_order, scale, scale_e, xr_detect, nexp=0):
    Zn = _c128(Zn_path)
    L = Zn.shape[0]
    idx, rx, rxe = _i32(ref_index_xr), _c128(ref_xr), _i32(ref_xr_e)
    n_xr = 0 if idx is None else idx.shape[0]
    out = np.zeros(L, np.complex128)
    oe = np.zeros(L, np.int32)
    lib().fso_dzndc_path_mn(int(nexp or 0), c_vp(_p(Zn)), c_i64(L), c_i64(n_xr), c_vp(_p(idx)),
                            c_vp(_p(rx)), c_vp(_p(rxe)), c_i64(ref_div_iter),
                            c_i64(ref_order), c_dbl(scale), c_i32(scale_e),
                            int(bool(xr_detect)), c_vp(_p(out)), c_vp(_p(oe)))
    return out, (oe if xr_detect else None)


def dzndz_path_m2(Zn_path, ref_index_xr, ref_xr, ref_xr_e, ref_div_iter,
                  ref_order, xr_detect, nexp=0):
    Zn = _c128(Zn_path)
    L = Zn.shape[0]
    idx, rx, rxe = _i32(ref_index_xr), _c128(ref_xr), _i32(ref_xr_e)
    n_xr = 0 if idx is None else idx.shape[0]
    out = np.zeros(L + 1, np.complex128)
    oe = np.zeros(L + 1, np.int32)
    lib().fso_dzndz_path_mn(int(nexp or 0), c_vp(_p(Zn)), c_i64(L), c_i64(n_xr), c_vp(_p(idx)),
                            c_vp(_p(rx)), c_vp(_p(rxe)), c_i64(ref_div_iter),
                            c_i64(ref_order), int(bool(xr_detect)),
                            c_vp(_p(out)), c_vp(_p(oe)))
    return out, (oe if xr_detect else None)


def dzndc_path_bs(flavor, Zn_path, ref_index_xr, refx_xr, refx_xr_e, refy_xr,
                  refy_xr_e, ref_div_iter, ref_order, scale, scale_e,
                  xr_detect):
    Zn = _c128(Zn_path)
    L = Zn.shape[0]
    idx = _i32(ref_index_xr)
    n_xr = 0 if idx is None else idx.shape[0]
    rx, rxe, ry, rye = _f64(refx_xr), _i32(refx_xr_e), _f64(refy_xr), _i32(refy_xr_e)
    out = np.zeros((4, L), np.float64)
    oe = np.zeros((4, L), np.int32)
    lib().fso_dzndc_path_bs(int(flavor), c_vp(_p(Zn)), c_i64(L), c_i64(n_xr),
                            c_vp(_p(idx)), c_vp(_p(rx)), c_vp(_p(rxe)),
                            c_vp(_p(ry)), c_vp(_p(rye)), c_i64(ref_div_iter),
                            c_i64(ref_order), c_dbl(scale), c_i32(scale_e),
                            int(bool(xr_detect)), c_vp(_p(out)), c_vp(_p(oe)))
    return out, (oe if xr_detect else None)


# ---------------------------------------------------------------------------
# Xrange unit-test entry points
def xr_binop_c(op, a, ae, b, be):
    a, b, ae, be = _c128(a), _c128(b), _i32(ae), _i32(be)
    n = a.shape[0]
    out = np.zeros(n, np.complex128)
    oe = np.zeros(n, np.int32)
    lib().fso_xr_binop_c(int(op), c_i64(n), c_vp(_p(a)), c_vp(_p(ae)),
                         c_vp(_p(b)), c_vp(_p(be)), c_vp(_p(out)), c_vp(_p(oe)))
    return out, oe


def xr_binop_f(op, a, ae, b, be):
    a, b, ae, be = _f64(a), _f64(b), _i32(ae), _i32(be)
    n = a.shape[0]
    out = np.zeros(n, np.float64)
    oe = np.zeros(n, np.int32)
    lib().fso_xr_binop_f(int(op), c_i64(n), c_vp(_p(a)), c_vp(_p(ae)),
                         c_vp(_p(b)), c_vp(_p(be)), c_vp(_p(out)), c_vp(_p(oe)))
    return out, oe


def xr_compare_f(cmp, a, ae, b, be):
    a, b, ae, be = _f64(a), _f64(b), _i32(ae), _i32(be)
    n = a.shape[0]
    out = np.zeros(n, np.uint8)
    lib().fso_xr_compare_f(int(cmp), c_i64(n), c_vp(_p(a)), c_vp(_p(ae)),
                           c_vp(_p(b)), c_vp(_p(be)), c_vp(_p(out)))
    return out.astype(bool)


def xr_to_standard_c(a, ae):
    a, ae = _c128(a), _i32(ae)
    out = np.zeros(a.shape[0], np.complex128)
    lib().fso_xr_to_standard_c(c_i64(a.shape[0]), c_vp(_p(a)), c_vp(_p(ae)),
                               c_vp(_p(out)))
    return out


def xr_to_standard_f(a, ae):
    a, ae = _f64(a), _i32(ae)
    out = np.zeros(a.shape[0], np.float64)
    lib().fso_xr_to_standard_f(c_i64(a.shape[0]), c_vp(_p(a)), c_vp(_p(ae)),
                               c_vp(_p(out)))
    return out


def xr_normalize_c(a, ae):
    a, ae = _c128(a), _i32(ae)
    out = np.zeros(a.shape[0], np.complex128)
    oe = np.zeros(a.shape[0], np.int32)
    lib().fso_xr_normalize_c(c_i64(a.shape[0]), c_vp(_p(a)), c_vp(_p(ae)),
                             c_vp(_p(out)), c_vp(_p(oe)))
    return out, oe
