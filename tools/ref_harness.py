# -*- coding: utf-8 -*-
"""
ref_harness.py -- run the UNMODIFIED reference (GBillotey/Fractalshades numba
code imported from /root/reference/src) in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product, in `pytest -m gpu`, in
`smoke()` or in `bench.py` imports this module: /root/reference does not exist
on the GPU box.  It is used by `tools/gen_golden.py` to produce the committed
fixtures under tests/golden/ and by ad-hoc validation of the C oracle.

The reference's only compiled component (`mpmath_utils/FP_loop.pyx`, Cython
over gmpy2/MPFR) cannot be built here (no gmpy2, no MPFR headers), so a shim
module with the same 3-tuple contract is injected; the shim calls this repo's
native MPFR orbit (libfsb200_orbit.so, include/fsb200_orbit.h) and wraps the
result in the reference's own `Xrange_array` (recipe: SURVEY.md appendix C).
"""
import ctypes
import os
import sys
import types
import tempfile

import numpy as np

REF_SRC = os.environ.get("FS_REF_SRC", "/root/reference/src")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_fs = None


class _XR(ctypes.Structure):
    _fields_ = [("index", ctypes.c_int64), ("mx", ctypes.c_double),
                ("my", ctypes.c_double), ("ex", ctypes.c_int32),
                ("ey", ctypes.c_int32)]


def _orbit_lib():
    lib = ctypes.CDLL(os.path.join(REPO, "fractalshades_b200",
                                   "libfsb200_orbit.so"))
    common = [ctypes.c_int, ctypes.c_double, ctypes.c_char_p, ctypes.c_char_p,
              ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
              ctypes.POINTER(ctypes.c_int64)]
    lib.fsb_orbit_mandelbrot.restype = ctypes.c_int64
    lib.fsb_orbit_mandelbrot.argtypes = [ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_uint32] + common
    lib.fsb_orbit_burning_ship.restype = ctypes.c_int64
    lib.fsb_orbit_burning_ship.argtypes = [ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_int] + common
    return lib


def _make_shim(fsx):
    """ Module standing in for fractalshades.mpmath_utils.FP_loop """
    lib = _orbit_lib()
    shim = types.ModuleType("fractalshades.mpmath_utils.FP_loop")
    XR_CAP = 1 << 20

    def _call(fn, orbit, need_xr, max_iter, extra, M, sx, sy, prec):
        assert orbit.dtype == np.float64 and orbit.shape[0] == 2 * (max_iter + 1)
        orbit = np.ascontiguousarray(orbit)
        buf = (_XR * XR_CAP)()
        cnt = ctypes.c_int64(0)
        i = fn(orbit.ctypes.data, max_iter, extra, int(bool(need_xr)), M, sx,
               sy, prec, buf, XR_CAP, ctypes.byref(cnt))
        if i < 0:
            raise RuntimeError(f"native orbit failed: {i}")
        return i, buf, cnt.value

    def perturbation_mandelbrot_FP_loop(orbit, need_Xrange, max_iter, M,
                                        seed_x, seed_y, seed_prec):
        return perturbation_mandelbrotN_FP_loop(
            orbit, need_Xrange, max_iter, 2, M, seed_x, seed_y, seed_prec)

    def perturbation_mandelbrotN_FP_loop(orbit, need_Xrange, max_iter,
                                         exponent, M, seed_x, seed_y,
                                         seed_prec):
        i, buf, n = _call(lib.fsb_orbit_mandelbrot, orbit, need_Xrange,
                          max_iter, exponent, M, seed_x, seed_y, seed_prec)
        xr = {}
        for k in range(n):
            e = buf[k]
            # FP_loop.pyx:424-441
            x_Xr = fsx.Xrange_array([e.mx], e.ex, np.complex128)
            y_Xr = fsx.Xrange_array([e.my], e.ey, np.complex128)
            xr[int(e.index)] = (x_Xr + 1j * y_Xr)
        return i, {}, xr

    def perturbation_nonholomorphic_FP_loop(orbit, need_Xrange, max_iter, M,
                                            seed_x, seed_y, seed_prec, kind):
        i, buf, n = _call(lib.fsb_orbit_burning_ship, orbit, need_Xrange,
                          max_iter, kind, M, seed_x, seed_y, seed_prec)
        xr = {}
        for k in range(n):
            e = buf[k]
            # FP_loop.pyx:443-454 (note: complex dtype in the reference)
            x_Xr = fsx.Xrange_array([e.mx], e.ex, np.complex128)
            y_Xr = fsx.Xrange_array([e.my], e.ey, np.complex128)
            xr[int(e.index)] = (x_Xr, y_Xr)
        return i, {}, xr

    # period / nucleus search: same contract as FP_loop.pyx:605-758, 900-1340
    lib.fsb_ball_method_mandelbrot.restype = ctypes.c_int64
    lib.fsb_ball_method_mandelbrot.argtypes = [
        ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p,
        ctypes.c_int64, ctypes.c_double]
    lib.fsb_find_nucleus_mandelbrot.restype = ctypes.c_int
    lib.fsb_find_nucleus_mandelbrot.argtypes = [
        ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64,
        ctypes.c_int64, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
        ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64]

    def perturbation_mandelbrot_ball_method(seed_x, seed_y, seed_prec, seed_px,
                                            maxiter, M_divergence):
        return int(lib.fsb_ball_method_mandelbrot(seed_x, seed_y, seed_prec, seed_px,
                                                  maxiter, M_divergence))

    def _newton(any_nucleus):
        def impl(seed_x, seed_y, seed_prec, order, max_newton, seed_eps_cv,
                 seed_eps_valid):
            import mpmath
            cap = int(seed_prec * 0.31) + 64
            bx = ctypes.create_string_buffer(cap)
            by = ctypes.create_string_buffer(cap)
            rc = lib.fsb_find_nucleus_mandelbrot(seed_x, seed_y, seed_prec, order,
                                                 max_newton, seed_eps_cv, seed_eps_valid,
                                                 any_nucleus, bx, by, cap)
            if rc != 1:
                return False, mpmath.mpc("nan", "nan")
            with mpmath.workprec(seed_prec):
                return True, mpmath.mpc(mpmath.mpf(bx.value.decode()),
                                        mpmath.mpf(by.value.decode()))
        return impl

    # burning-ship family and z^N + c: same contracts (FP_loop.pyx:2357-2755, 631-758,
    # 1159-1340), served by the native library as well
    cp = ctypes.c_char_p
    for name, first in (("burning_ship", ctypes.c_int), ("mandelbrot_n", ctypes.c_uint32)):
        fn = getattr(lib, "fsb_ball_method_" + name)
        fn.restype = ctypes.c_int64
        fn.argtypes = [first, cp, cp, ctypes.c_int64, cp, ctypes.c_int64, ctypes.c_double]
        fn = getattr(lib, "fsb_find_any_nucleus_" + name)
        fn.restype = ctypes.c_int
        fn.argtypes = [first, cp, cp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, cp, cp,
                       cp, cp, ctypes.c_int64]

    def _any(fn_name, first, seed_x, seed_y, seed_prec, order, max_newton, seed_eps_cv,
             seed_eps_valid):
        import mpmath
        cap = int(seed_prec * 0.31) + 64
        bx = ctypes.create_string_buffer(cap)
        by = ctypes.create_string_buffer(cap)
        rc = getattr(lib, fn_name)(first, seed_x, seed_y, seed_prec, order, max_newton,
                                   seed_eps_cv, seed_eps_valid, bx, by, cap)
        if rc != 1:
            return False, mpmath.mpc("nan", "nan")
        with mpmath.workprec(seed_prec):
            return True, mpmath.mpc(mpmath.mpf(bx.value.decode()), mpmath.mpf(by.value.decode()))

    def perturbation_nonholomorphic_ball_method(seed_x, seed_y, seed_prec, seed_px, maxiter,
                                                M_divergence, kind):
        return int(lib.fsb_ball_method_burning_ship(kind, seed_x, seed_y, seed_prec, seed_px,
                                                    maxiter, M_divergence))

    def perturbation_nonholomorphic_find_any_nucleus(seed_x, seed_y, seed_prec, order,
                                                     max_newton, seed_eps_cv, seed_eps_valid,
                                                     kind):
        return _any("fsb_find_any_nucleus_burning_ship", kind, seed_x, seed_y, seed_prec, order,
                    max_newton, seed_eps_cv, seed_eps_valid)

    def perturbation_mandelbrotN_ball_method(seed_x, seed_y, seed_prec, seed_px, exponent,
                                             maxiter, M_divergence):
        return int(lib.fsb_ball_method_mandelbrot_n(exponent, seed_x, seed_y, seed_prec, seed_px,
                                                    maxiter, M_divergence))

    def perturbation_mandelbrotN_find_any_nucleus(seed_x, seed_y, seed_prec, exponent, order,
                                                  max_newton, seed_eps_cv, seed_eps_valid):
        return _any("fsb_find_any_nucleus_mandelbrot_n", exponent, seed_x, seed_y, seed_prec,
                    order, max_newton, seed_eps_cv, seed_eps_valid)

    shim.perturbation_nonholomorphic_ball_method = perturbation_nonholomorphic_ball_method
    shim.perturbation_nonholomorphic_find_any_nucleus = perturbation_nonholomorphic_find_any_nucleus
    shim.perturbation_mandelbrotN_ball_method = perturbation_mandelbrotN_ball_method
    shim.perturbation_mandelbrotN_find_any_nucleus = perturbation_mandelbrotN_find_any_nucleus
    shim.perturbation_mandelbrot_ball_method = perturbation_mandelbrot_ball_method
    shim.perturbation_mandelbrot_find_nucleus = _newton(0)
    shim.perturbation_mandelbrot_find_any_nucleus = _newton(1)
    shim.perturbation_mandelbrot_FP_loop = perturbation_mandelbrot_FP_loop
    shim.perturbation_mandelbrotN_FP_loop = perturbation_mandelbrotN_FP_loop
    shim.perturbation_nonholomorphic_FP_loop = perturbation_nonholomorphic_FP_loop
    return shim


def load_reference():
    """ Import the reference with the FP_loop shim; returns the package. """
    global _fs
    if _fs is not None:
        return _fs
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"reference sources not found at {REF_SRC}")
    sys.path.insert(0, REF_SRC)
    if os.environ.get("FS_REF_STRICT", "0") == "1":
        # Validation mode: compile the reference WITHOUT fastmath (and without
        # its on-disk cache) so that it becomes an IEEE-strict sequence the
        # oracle must then reproduce bit for bit.  The reference sources are
        # untouched; only numba.njit's keyword arguments are filtered.
        import numba
        _orig_njit = numba.njit

        def _strict_njit(*args, **kwargs):
            kwargs["fastmath"] = False
            kwargs["cache"] = False
            return _orig_njit(*args, **kwargs)
        numba.njit = _strict_njit
    import fractalshades.numpy_utils.xrange as fsx
    import fractalshades.mpmath_utils as mu
    shim = _make_shim(fsx)
    sys.modules["fractalshades.mpmath_utils.FP_loop"] = shim
    mu.FP_loop = shim
    import fractalshades as fs
    import fractalshades.models  # noqa
    fs.settings.no_newton = True
    fs.settings.enable_multithreading = True
    import logging
    logging.getLogger("fractalshades").setLevel(logging.ERROR)
    _fs = fs
    return fs


def xr_to_pairs(arr):
    """ Xrange_array (any shape) -> (mantissa ndarray, exp int32 ndarray) """
    a = np.asarray(arr)
    return np.array(a["mantissa"]), np.array(a["exp"], dtype=np.int32)


def make_fractal(case, workdir):
    fs = load_reference()
    import fractalshades.models as fsm
    kind = case["kind"]
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from cases import make_projection
    proj = make_projection(fs.projection, case.get("proj"))
    if kind == "std_M2":
        if "exponent" in case.get("init", {}):
            f = fsm.Mandelbrot_N(workdir, **case["init"])
        else:
            f = fsm.Mandelbrot(workdir)
        f.zoom(x=case["x"], y=case["y"], dx=case["dx"], nx=case["nx"],
               xy_ratio=case.get("xy_ratio", 1.0),
               theta_deg=case.get("theta_deg", 0.), projection=proj,
               **case.get("skew", {}))
        f.calc_std_div(calc_name="c", subset=None, **case["calc"])
    elif kind == "std_BS":
        f = fsm.Burning_ship(workdir, **case.get("init", {}))
        f.zoom(x=case["x"], y=case["y"], dx=case["dx"], nx=case["nx"],
               xy_ratio=case.get("xy_ratio", 1.0),
               theta_deg=case.get("theta_deg", 0.), projection=proj,
               **case.get("skew", {}))
        f.calc_std_div(calc_name="c", subset=None, **case["calc"])
    elif kind == "perturb_M2":
        if "exponent" in case.get("init", {}):
            f = fsm.Perturbation_mandelbrot_N(workdir, **case["init"])
        else:
            f = fsm.Perturbation_mandelbrot(workdir)
        f.zoom(precision=case["precision"], x=case["x"], y=case["y"],
               dx=case["dx"], nx=case["nx"],
               xy_ratio=case.get("xy_ratio", 1.0),
               theta_deg=case.get("theta_deg", 0.), projection=proj,
               **case.get("skew", {}))
        f.calc_std_div(calc_name="c", subset=None, **case["calc"])
    elif kind == "perturb_BS":
        f = fsm.Perturbation_burning_ship(workdir, **case.get("init", {}))
        f.zoom(precision=case["precision"], x=case["x"], y=case["y"],
               dx=case["dx"], nx=case["nx"],
               xy_ratio=case.get("xy_ratio", 1.0),
               theta_deg=case.get("theta_deg", 0.), projection=proj,
               **case.get("skew", {}))
        f.calc_std_div(calc_name="c", subset=None, **case["calc"])
    else:
        raise ValueError(kind)
    step = (case.get("proj") or {}).get("step")
    if step is not None:
        # one step of the stepped exponential zoom (core.py:891-963): the
        # projection is told the step bounds, then the frame tables are rebuilt
        proj.set_exp_zoom_step(*step)
        data = f._calc_data["c"]
        data["cycle_indep_args"] = f.reset_bla_tree(data["cycle_indep_args"])
    return f


def run_case(case, workdir=None, keep_tables=True):
    """
    Run one case through the reference.  Returns a dict with the per-pixel
    outputs in chunk-rank (memmap) order and, for perturbation cases, the
    frame tables (`cycle_indep_args`).
    """
    own = workdir is None
    if own:
        tmp = tempfile.TemporaryDirectory()
        workdir = tmp.name
    fs = load_reference()
    # cases flagged newton=True run the reference's default flow (ball method +
    # Newton descent -> periodic reference); all others use the image centre
    fs.settings.no_newton = not case.get("newton", False)
    try:
        f = make_fractal(case, workdir)
        f.calc_raw("c")
    finally:
        fs.settings.no_newton = True
    out = {}
    for key in ("Z", "U", "stop_reason", "stop_iter"):
        out[key] = np.array(f.get_data_memmap("c", key, mode="r"))
    cpix = []
    for chunk_slice in f.chunk_slices():
        cpix.append(np.ravel(f.chunk_pixel_pos(chunk_slice, False, None)))
    out["c_pix"] = np.concatenate(cpix)
    out["nx"], out["ny"] = f.nx, f.ny
    out["lin_mat"] = np.array(f.lin_mat)
    if keep_tables:
        out["indep"] = f._calc_data["c"]["cycle_indep_args"]
    out["fractal"] = f
    if own:
        out["_tmp"] = tmp
    return out
