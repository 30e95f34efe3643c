# -*- coding: utf-8 -*-
"""
ctypes binding of the two native libraries of this package:

  libfsb200.so        CUDA sm_100a kernels + C ABI   (include/fsb200.h)
  libfsb200_orbit.so  MPFR reference orbit, host only (include/fsb200_orbit.h)

There is NO CPU fallback for the pixel path: if libfsb200.so is missing, or no
CUDA device is visible, every compute call raises RuntimeError.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p


class FsbStats(ctypes.Structure):
    _fields_ = [("kernel_ms", c_dbl), ("h2d_ms", c_dbl), ("d2h_ms", c_dbl),
                ("n_iter_exec", c_i64), ("n_bla_steps", c_i64),
                ("n_rebase", c_i64), ("sum_stop_iter", c_i64),
                ("n_launches", c_i64), ("n_iter_fast", c_i64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class FsbProjDesc(ctypes.Structure):
    """ fsb_proj_desc: projection + dz/dc modifier parameters (all zero =
    Cartesian, no modifier) """
    _fields_ = [("kind", c_i32), ("dzndc_modifier", c_i32), ("hmoy", c_dbl),
                ("pix_to_ht", c_dbl * 2), ("mod_param", c_dbl)]


class FsbStdDesc(ctypes.Structure):
    _fields_ = [("model", c_i32), ("flavor", c_i32), ("center_re", c_dbl),
                ("center_im", c_dbl), ("dx", c_dbl), ("lin_mat", c_dbl * 4),
                ("max_iter", c_i64), ("M_divergence_sq", c_dbl),
                ("epsilon_stationnary_sq", c_dbl), ("calc_d2zndc2", c_i32),
                ("calc_orbit", c_i32), ("backshift", c_i64),
                ("proj", FsbProjDesc), ("nexp", c_i32), ("_pad", c_i32)]


class FsbFrameDesc(ctypes.Structure):
    _fields_ = [
        ("model", c_i32), ("flavor", c_i32), ("L", c_i64), ("Zn_path", c_vp),
        ("n_xr", c_i64), ("ref_index_xr", c_vp), ("ref_xr", c_vp),
        ("ref_xr_e", c_vp), ("refy_xr", c_vp), ("refy_xr_e", c_vp),
        ("ref_div_iter", c_i64), ("ref_order", c_i64), ("drift", c_dbl * 2),
        ("drift_e", c_i32 * 2), ("lin_scale", c_dbl), ("lin_scale_e", c_i32),
        ("_pad0", c_i32), ("lin_mat", c_dbl * 4), ("kc", c_dbl), ("kc_e", c_i32),
        ("_pad1", c_i32), ("scale_deriv", c_dbl), ("scale_deriv_e", c_i32),
        ("_pad2", c_i32), ("xr_detect", c_i32), ("bla_activated", c_i32),
        ("calc_dzndc", c_i32), ("calc_dzndz", c_i32), ("calc_orbit", c_i32),
        ("_pad3", c_i32), ("backshift", c_i64), ("max_iter", c_i64),
        ("M_divergence_sq", c_dbl), ("epsilon_stationnary_sq", c_dbl),
        ("BLA_eps", c_dbl), ("dZndc", c_vp), ("dZndc_e", c_vp), ("dZndz", c_vp),
        ("dZndz_e", c_vp), ("M_bla", c_vp), ("r_bla", c_vp), ("bla_len", c_i64),
        ("stages_bla", c_i32), ("nexp", c_i32), ("proj", FsbProjDesc),
    ]


class OrbitXr(ctypes.Structure):
    _fields_ = [("index", c_i64), ("mx", c_dbl), ("my", c_dbl),
                ("ex", c_i32), ("ey", c_i32)]


FSB_MODEL_M2 = 0
FSB_MODEL_BS = 1
FSB_PROJ_CARTESIAN, FSB_PROJ_EXPMAP = 0, 1
FSB_DZNDC_MOD_NONE, FSB_DZNDC_MOD_EXPMAP, FSB_DZNDC_MOD_SEAM = 0, 1, 2

# Every symbol declared in include/fsb200.h (checked by the CPU test-suite)
CUDA_SYMBOLS = [
    "fsb_device_count", "fsb_init", "fsb_shutdown", "fsb_last_error",
    "fsb_build_info", "fsb_device_info", "fsb_host_alloc", "fsb_host_free",
    "fsb_dev_alloc", "fsb_dev_free", "fsb_memcpy_h2d", "fsb_memcpy_d2h",
    "fsb_dev_memset", "fsb_flush_l2", "fsb_std_nz", "fsb_std_run",
    "fsb_std_run_device", "fsb_std_run_tiles", "fsb_std_run_tiles_device",
    "fsb_frame_run_tiles", "fsb_frame_run_tiles_device", "fsb_frame_create", "fsb_frame_destroy",
    "fsb_frame_nz", "fsb_frame_bla_len", "fsb_frame_stages_bla",
    "fsb_frame_setup_ms", "fsb_frame_get_bla", "fsb_frame_get_dzndc",
    "fsb_frame_get_dzndz", "fsb_frame_run", "fsb_frame_run_device",
    "fsb_frame_run_pp", "fsb_postproc_run", "fsb_postproc_run_device",
    "fsb_xr_binop_c", "fsb_xr_to_standard_c", "fsb_hypot_test",
    "fsb_fp64_peak_tflops", "fsb_proj_apply",
    "fsb_std_run_grid", "fsb_frame_run_grid", "fsb_frame_run_grid_pp",
    "fsb_postproc_ext_run", "fsb_postproc_ext_run_device", "fsb_frame_run_grid_pp_ext",
    "fsb_postproc_run_proj", "fsb_postproc_run_proj_device",
]
ORBIT_SYMBOLS = ["fsb_orbit_mandelbrot", "fsb_orbit_burning_ship",
                 "fsb_ball_method_mandelbrot", "fsb_find_nucleus_mandelbrot",
                 "fsb_ball_method_burning_ship", "fsb_find_any_nucleus_burning_ship",
                 "fsb_ball_method_mandelbrot_n", "fsb_find_any_nucleus_mandelbrot_n"]

_libs = {}


def lib_path(strict=False):
    override = os.environ.get("FSB200_LIB")          # kernel A/B experiments
    if override:
        return override
    if not strict and os.environ.get("FSB200_LIB_DEFAULT"):     # default build only
        return os.environ["FSB200_LIB_DEFAULT"]
    return os.path.join(_PKG, "libfsb200_strict.so" if strict else "libfsb200.so")


def orbit_lib_path():
    return os.path.join(_PKG, "libfsb200_orbit.so")


def _declare(lib):
    lib.fsb_last_error.restype = ctypes.c_char_p
    lib.fsb_build_info.restype = ctypes.c_char_p
    lib.fsb_host_alloc.restype = c_vp
    lib.fsb_host_alloc.argtypes = [c_i64]
    lib.fsb_host_free.argtypes = [c_vp]
    lib.fsb_dev_alloc.restype = c_vp
    lib.fsb_dev_alloc.argtypes = [c_i64]
    lib.fsb_dev_free.argtypes = [c_vp]
    lib.fsb_memcpy_h2d.argtypes = [c_vp, c_vp, c_i64]
    lib.fsb_memcpy_d2h.argtypes = [c_vp, c_vp, c_i64]
    lib.fsb_dev_memset.argtypes = [c_vp, ctypes.c_int, c_i64]
    lib.fsb_device_info.argtypes = [ctypes.c_char_p, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(c_i64)]
    lib.fsb_std_nz.argtypes = [ctypes.POINTER(FsbStdDesc)]
    lib.fsb_std_run.argtypes = [ctypes.POINTER(FsbStdDesc), c_i64, c_vp, c_vp,
                                c_vp, c_vp, c_vp, ctypes.POINTER(FsbStats)]
    lib.fsb_std_run_device.argtypes = [ctypes.POINTER(FsbStdDesc), c_i64, c_vp,
                                       c_vp, c_vp, c_vp,
                                       ctypes.POINTER(FsbStats)]
    lib.fsb_std_run_tiles.argtypes = [ctypes.POINTER(FsbStdDesc), c_i32, c_vp, c_vp,
                                      c_vp, c_vp, c_vp, c_vp, c_vp,
                                      ctypes.POINTER(FsbStats)]
    lib.fsb_std_run_tiles_device.argtypes = [ctypes.POINTER(FsbStdDesc), c_i32, c_vp,
                                             c_vp, c_vp, c_vp, c_vp, c_vp,
                                             ctypes.POINTER(FsbStats)]
    lib.fsb_frame_run_tiles.argtypes = [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, ctypes.POINTER(FsbStats)]
    lib.fsb_frame_run_tiles_device.argtypes = [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp,
                                               c_vp, c_vp, c_vp,
                                               ctypes.POINTER(FsbStats)]
    lib.fsb_frame_create.argtypes = [ctypes.POINTER(FsbFrameDesc),
                                     ctypes.POINTER(c_vp)]
    lib.fsb_frame_destroy.argtypes = [c_vp]
    lib.fsb_frame_nz.argtypes = [c_vp]
    lib.fsb_frame_bla_len.argtypes = [c_vp]
    lib.fsb_frame_bla_len.restype = c_i64
    lib.fsb_frame_stages_bla.argtypes = [c_vp]
    lib.fsb_frame_setup_ms.argtypes = [c_vp, ctypes.c_int]
    lib.fsb_frame_setup_ms.restype = c_dbl
    lib.fsb_frame_get_bla.argtypes = [c_vp, c_vp, c_vp]
    lib.fsb_frame_get_dzndc.argtypes = [c_vp, c_vp, c_vp]
    lib.fsb_frame_get_dzndz.argtypes = [c_vp, c_vp, c_vp]
    lib.fsb_frame_run.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, ctypes.POINTER(FsbStats)]
    lib.fsb_frame_run_device.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, ctypes.POINTER(FsbStats)]
    lib.fsb_xr_binop_c.argtypes = [ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp,
                                   c_vp, c_vp]
    lib.fsb_xr_to_standard_c.argtypes = [c_i64, c_vp, c_vp, c_vp]
    lib.fsb_hypot_test.argtypes = [c_i64, c_vp, c_vp, c_vp]
    if hasattr(lib, "fsb_proj_apply"):     # absent from older A/B builds (FSB200_LIB)
        lib.fsb_proj_apply.argtypes = [ctypes.POINTER(FsbProjDesc), c_i64, c_vp, c_vp, c_vp]
    if hasattr(lib, "fsb_frame_run_grid"):
        lib.fsb_std_run_grid.argtypes = [ctypes.POINTER(FsbStdDesc), c_i32, c_vp, c_vp,
                                         c_vp, c_vp, c_vp, c_vp, c_vp,
                                         ctypes.POINTER(FsbStats)]
        lib.fsb_frame_run_grid.argtypes = [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                           c_vp, c_vp, c_vp, ctypes.POINTER(FsbStats)]
    lib.fsb_fp64_peak_tflops.argtypes = [ctypes.c_int]
    lib.fsb_fp64_peak_tflops.restype = c_dbl
    return lib


def load_cuda_lib(strict=False):
    """ dlopen the CUDA library (does not touch the device). """
    key = "strict" if strict else "default"
    if key not in _libs:
        path = lib_path(strict)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` "
                "(the pixel path has no CPU fallback)")
        _libs[key] = _declare(ctypes.CDLL(path))
    return _libs[key]


def cuda_lib(strict=None):
    """ The CUDA library bound to a device; raises if there is no GPU. """
    if strict is None:
        from . import settings
        strict = bool(settings.strict_ieee)
    lib = load_cuda_lib(strict)
    key = "init_strict" if strict else "init_default"
    if key not in _libs:
        dev = int(os.environ.get("LOCAL_RANK", os.environ.get("FSB200_DEVICE", "0")))
        n = lib.fsb_device_count()
        if n <= 0:
            raise RuntimeError(
                "fractalshades_b200: no CUDA device available "
                f"({lib.fsb_last_error().decode()}); no CPU fallback exists")
        check(lib, lib.fsb_init(dev % n))
        _libs[key] = True
    return lib


def check(lib, rc):
    if rc < 0:
        raise RuntimeError("libfsb200: " + lib.fsb_last_error().decode())
    return rc


def load_orbit_lib():
    if "orbit" not in _libs:
        path = orbit_lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
        lib = ctypes.CDLL(path)
        common = [ctypes.c_int, c_dbl, ctypes.c_char_p, ctypes.c_char_p, c_i64,
                  c_vp, c_i64, ctypes.POINTER(c_i64)]
        lib.fsb_orbit_mandelbrot.restype = c_i64
        lib.fsb_orbit_mandelbrot.argtypes = [c_vp, c_i64, ctypes.c_uint32] + common
        lib.fsb_orbit_burning_ship.restype = c_i64
        lib.fsb_orbit_burning_ship.argtypes = [c_vp, c_i64, ctypes.c_int] + common
        lib.fsb_ball_method_mandelbrot.restype = c_i64
        lib.fsb_ball_method_mandelbrot.argtypes = [
            ctypes.c_char_p, ctypes.c_char_p, c_i64, ctypes.c_char_p, c_i64, c_dbl]
        lib.fsb_find_nucleus_mandelbrot.restype = ctypes.c_int
        lib.fsb_find_nucleus_mandelbrot.argtypes = [
            ctypes.c_char_p, ctypes.c_char_p, c_i64, c_i64, c_i64, ctypes.c_char_p,
            ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, c_i64]
        cp = ctypes.c_char_p
        for name, first in (("burning_ship", ctypes.c_int), ("mandelbrot_n", ctypes.c_uint32)):
            fn = getattr(lib, "fsb_ball_method_" + name)
            fn.restype = c_i64
            fn.argtypes = [first, cp, cp, c_i64, cp, c_i64, c_dbl]
            fn = getattr(lib, "fsb_find_any_nucleus_" + name)
            fn.restype = ctypes.c_int
            fn.argtypes = [first, cp, cp, c_i64, c_i64, c_i64, cp, cp, cp, cp, c_i64]
        _libs["orbit"] = lib
    return _libs["orbit"]


def newton_call(fn_name, first, c, order, eps_pixel, max_newton, eps_cv):
    """ shared wrapper of the fsb_find_any_nucleus_* entry points (the argument
    handling of the reference's model methods, e.g. burning_ship.py:1203-1232) """
    import mpmath
    if order is None:
        raise ValueError("order shall be defined for Newton method")
    seed_prec = mpmath.mp.prec
    if max_newton is None:
        max_newton = 80
    if eps_cv is None:
        eps_cv = mpmath.mpf(val=(2, -seed_prec))
    cap = int(seed_prec * 0.31) + 64
    bx, by = ctypes.create_string_buffer(cap), ctypes.create_string_buffer(cap)
    rc = getattr(load_orbit_lib(), fn_name)(
        first, str(c.real).encode("utf8"), str(c.imag).encode("utf8"), seed_prec, int(order),
        int(max_newton), str(eps_cv).encode("utf8"), str(eps_pixel).encode("utf8"), bx, by, cap)
    if rc < 0:
        raise RuntimeError(f"{fn_name} failed ({rc})")
    if rc == 0:
        return False, mpmath.mpc("nan", "nan")
    return True, mpmath.mpc(mpmath.mpf(bx.value.decode()), mpmath.mpf(by.value.decode()))


def ptr(a):
    return None if a is None else a.ctypes.data


def pinned_empty(shape, dtype, strict=None):
    """ numpy array backed by pinned (page-locked) host memory. """
    lib = cuda_lib(strict)
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib.fsb_host_alloc(max(n, 1))
    if not p:
        raise RuntimeError("libfsb200: " + lib.fsb_last_error().decode())
    buf = (ctypes.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _pinned_registry[arr.ctypes.data] = (lib, p)
    return arr


_pinned_registry = {}


def pinned_free(arr):
    ent = _pinned_registry.pop(arr.ctypes.data, None)
    if ent is not None:
        ent[0].fsb_host_free(ent[1])
