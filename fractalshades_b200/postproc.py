# -*- coding: utf-8 -*-
"""
Post-processing of the raw fields on the GPU.

The reference computes its post-processed fields tile by tile in numpy from the
raw memmaps (`postproc.py`): `Continuous_iter_pp` (:243-406, 1001-1009),
`DEM_pp` (:640-731), `DEM_normal_pp` (:520-628, kind "potential").  Here the
same three fields are evaluated by one elementwise kernel (`k_postproc`,
include/fsb200.h `fsb_postproc_*`) either

  * fused with the pixel kernels of a perturbation frame (`frame_fields`): the
    raw Z / U planes never leave the device, only 4 bytes per point and field
    come back (the reference's default `postproc_dtype` is float32), or
  * stand-alone on raw arrays that are already on the host (`fields_from_raw`).

Values are meaningful where `stop_reason == 1` (escaped points), as in the
reference.  There is no CPU path: the numpy restatement used to check the
kernel lives in tests/test_gpu_postproc.py.
"""
import ctypes

import numpy as np

from . import _native
from .core import tile_shape_arrays

FIELDS = ("cont_iter", "DEM", "normal_x", "normal_y")


class FsbPostprocDesc(ctypes.Structure):
    _fields_ = [("holomorphic", ctypes.c_int32), ("row_zn", ctypes.c_int32),
                ("row_dzndc", ctypes.c_int32), ("has_skew", ctypes.c_int32),
                ("potential_d", ctypes.c_double), ("potential_a_d", ctypes.c_double),
                ("potential_M", ctypes.c_double), ("floor_iter", ctypes.c_double),
                ("px_snap", ctypes.c_double), ("skew", ctypes.c_double * 4),
                ("out_f64", ctypes.c_int32), ("_pad", ctypes.c_int32)]


def _declare(lib):
    if getattr(lib, "_pp_declared", False):
        return lib
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    D = ctypes.POINTER(FsbPostprocDesc)
    lib.fsb_frame_run_pp.argtypes = [vp, i32, vp, vp, i64, vp, D, vp, vp, vp, vp, vp, vp, vp,
                                     ctypes.POINTER(_native.FsbStats)]
    lib.fsb_frame_run_grid_pp.argtypes = [vp, i32, vp, vp, vp, D, vp, vp, vp, vp, vp, vp, vp,
                                          ctypes.POINTER(_native.FsbStats)]
    lib.fsb_postproc_run.argtypes = [D, i64, i32, vp, vp, vp, vp, vp, vp]
    lib.fsb_postproc_run_device.argtypes = [D, i64, i32, vp, vp, vp, vp, vp, vp]
    lib._pp_declared = True
    return lib


def make_desc(fractal, calc_name, floor_iter=0, px_snap=None, dtype=np.float32):
    """ Description of the post-processing for one calculation of `fractal`:
    rows of Z from the calculation's field codes, potential from the model
    (`potential_kind` must be "infinity": the divergent models). """
    state = fractal._calc_data[calc_name]["state"]
    codes = list(state.codes[0])
    if getattr(fractal, "potential_kind", "infinity") != "infinity":
        raise NotImplementedError("only the 'infinity' potential is built")
    d = FsbPostprocDesc()
    holo = np.dtype(state.complex_type) == np.complex128
    d.holomorphic = int(holo)
    d.row_zn = codes.index("zn" if holo else "xn")
    key = "dzndc" if holo else "dxnda"
    d.row_dzndc = codes.index(key) if key in codes else -1
    d.potential_d = float(fractal.potential_d)
    d.potential_a_d = float(fractal.potential_a_d)
    d.potential_M = float(getattr(state, "potential_M", fractal.potential_M_cutoff))
    d.floor_iter = float(floor_iter)
    d.px_snap = -1. if px_snap is None else float(px_snap)
    skew = getattr(fractal, "skew", None)
    d.has_skew = int(skew is not None)
    if skew is not None:
        for k, v in enumerate(np.asarray(skew, dtype=np.float64).ravel()):
            d.skew[k] = float(v)
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("postproc dtype must be float32 or float64")
    d.out_f64 = int(dtype == np.dtype(np.float64))
    return d, dtype, len(codes)


def _check_projection(fractal, names):
    """ The reference scales / rotates dzndc by the projection's derivative
    before DEM and normals (Postproc.get_dzndc -> apply_df / apply_dfBS with
    proj.df, postproc.py:184-206); k_postproc does not: refused, never
    approximated. """
    from . import projection as _projection
    if not (set(names) & {"DEM", "normal_x", "normal_y"}):
        return
    proj = fractal.projection
    if not (type(proj) is _projection.Cartesian and getattr(proj, "expmap_seam", None) is None):
        raise NotImplementedError(
            "GPU DEM / normal post-processing is defined for the plain Cartesian "
            f"projection only (got {type(proj).__name__}): the projection derivative of "
            "Postproc.get_dzndc is not applied by k_postproc")


def _outputs(fields, npts, dtype, have_deriv):
    want = set(fields)
    if "normal" in want:
        want |= {"normal_x", "normal_y"}
        want.discard("normal")
    unknown = want - set(FIELDS)
    if unknown:
        raise ValueError(f"unknown post-processing fields {sorted(unknown)}")
    if (want & {"DEM", "normal_x", "normal_y"}) and not have_deriv:
        raise ValueError("DEM / normal need the derivative fields (calc_dzndc / calc_hessian)")
    if ("normal_x" in want) != ("normal_y" in want):
        want |= {"normal_x", "normal_y"}
    return {k: np.empty(npts, dtype) for k in FIELDS if k in want}


def fields_from_raw(fractal, calc_name, Z, stop_iter, fields=("cont_iter", "DEM", "normal"),
                    floor_iter=0, px_snap=None, dtype=np.float32):
    """ Stand-alone: raw (n_fields, npts) arrays on the host -> dict of fields. """
    lib = _declare(_native.cuda_lib())
    d, dtype, n_rows = make_desc(fractal, calc_name, floor_iter, px_snap, dtype)
    Z = np.ascontiguousarray(Z)
    si = np.ascontiguousarray(np.ravel(stop_iter), dtype=np.int32)
    npts = Z.shape[1]
    out = _outputs(fields, npts, dtype, d.row_dzndc >= 0)
    _check_projection(fractal, out)
    rc = lib.fsb_postproc_run(ctypes.byref(d), npts, Z.shape[0], _native.ptr(Z), _native.ptr(si),
                              *[_native.ptr(out.get(k)) for k in FIELDS])
    _native.check(lib, rc)
    return out


_STAGING = {}


def _frame_staging(fractal, names, dtype, tiles=None):
    """ Per frame geometry, kept on the fractal: the per-tile pixel axes (they
    depend on nx, ny, xy_ratio only; the pixel grid itself is built on the
    device) and one page-locked output array per field. """
    tiles = list(fractal.chunk_slices()) if tiles is None else list(tiles)
    key = (fractal.nx, fractal.ny, fractal.xy_ratio, settings_chunk(), tuple(tiles))
    # one staging set per process (the frames of a zoom movie share it; page-locked
    # allocations of a few hundred MB cost more than an 8K frame's kernels)
    st = _STAGING.get("cur")
    if st is None or st["key"] != key:
        if st is not None:
            for a in st["bufs"].values():
                _native.pinned_free(a)
        from .core import TileAxes
        ta = TileAxes(fractal, tiles)
        st = {"key": key, "npts": ta.npts, "tw": ta.tw, "th": ta.th, "axes": ta.axes,
              "bufs": {}}
        _STAGING["cur"] = st
    for name, dt in names:
        cur = st["bufs"].get(name)
        if cur is None or cur.dtype != np.dtype(dt):
            if cur is not None:
                _native.pinned_free(cur)
            st["bufs"][name] = _native.pinned_empty((st["npts"],), dt)
    return st


def settings_chunk():
    from . import settings
    return settings.chunk_size


def frame_fields(fractal, calc_name, fields=("cont_iter", "DEM", "normal"), floor_iter=0,
                 px_snap=None, dtype=np.float32, want_stop_iter=False, copy=True, tiles=None):
    """ Fused: pixel kernels + post-processing of one whole perturbation frame
    (or of the listed tiles: `tiles` = chunk slices, e.g. one rank's share of
    the frame).  Returns (dict of tile-ordered 1-D fields incl. "stop_reason",
    stats).  Use `to_image` for the (ny, nx) arrays of a whole frame.
    copy=False returns views of the page-locked staging buffers (overwritten
    by the next call). """
    lib = _declare(_native.cuda_lib())
    indep = fractal._calc_data[calc_name]["cycle_indep_args"]
    if indep[0] != "perturb":
        raise NotImplementedError("fused post-processing: perturbation frames "
                                  "(use fields_from_raw for the standard models)")
    frame, interrupted = indep[1], indep[2]
    d, dtype, _ = make_desc(fractal, calc_name, floor_iter, px_snap, dtype)
    names = list(_outputs(fields, 0, dtype, d.row_dzndc >= 0))
    _check_projection(fractal, names)
    want = [(k, dtype) for k in names] + [("stop_reason", np.int8)]
    if want_stop_iter:
        want.append(("stop_iter", np.int32))
    st = _frame_staging(fractal, want, dtype, tiles)
    b, npts = st["bufs"], st["npts"]
    out = {k: b[k] for k, _ in want}
    stats = _native.FsbStats()
    rc = lib.fsb_frame_run_grid_pp(frame.ptr, st["tw"].shape[0], _native.ptr(st["tw"]),
                                   _native.ptr(st["th"]), _native.ptr(st["axes"]),
                                   ctypes.byref(d), *[_native.ptr(out.get(k)) for k in FIELDS],
                                   _native.ptr(out["stop_reason"]),
                                   _native.ptr(out.get("stop_iter")),
                                   _native.ptr(interrupted), stats)
    _native.check(lib, rc)
    if rc != 0:
        raise RuntimeError("frame interrupted")
    if copy:
        out = {k: np.array(v) for k, v in out.items()}
    return out, stats.as_dict()


def to_image(fractal, arr1d):
    """ tile-ordered 1-D field (chunk-rank order, row-major inside a tile,
    core.py:2362-2472) -> (ny, nx) array, row 0 = top of the image """
    img = np.empty((fractal.ny, fractal.nx), arr1d.dtype)
    off = 0
    for (ix, ixx, iy, iyy) in fractal.chunk_slices():
        w, h = ixx - ix, iyy - iy
        img[iy:iyy, ix:ixx] = arr1d[off:off + w * h].reshape(h, w)
        off += w * h
    return img
