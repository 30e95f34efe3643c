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

`Fieldlines_pp` (:409-531, 1038-1160) and the two per-pixel factors of
`Blinn_lighting.partial_shade` (colors/layers.py:865-903: Lambert and specular
coefficient of each light source, on the normal of `DEM_normal_pp` scaled by
sin(max_slope) as `Color_layer.apply_shade` :493-507 does) come from a second
elementwise kernel (`k_postproc_ext`, `fsb_postproc_ext_*`) over the same raw
rows, behind the same two entry points (`fieldlines=`, `lighting=`).

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
                ("out_f64", ctypes.c_int32), ("df_kind", ctypes.c_int32),
                ("df_k", ctypes.c_double * 2)]


PP_MAX_FL, PP_MAX_LIGHTS = 32, 4


class FsbPostprocExt(ctypes.Structure):
    _fields_ = [("fl_n_iter", ctypes.c_int32), ("fl_row_orbit", ctypes.c_int32),
                ("fl_backshift", ctypes.c_int32), ("fl_model", ctypes.c_int32),
                ("fl_k", ctypes.c_double * PP_MAX_FL), ("fl_phi", ctypes.c_double * PP_MAX_FL),
                ("c_center", ctypes.c_double * 2), ("c_scale", ctypes.c_double),
                ("c_lin_mat", ctypes.c_double * 4),
                ("n_lights", ctypes.c_int32), ("proj_kind", ctypes.c_int32),
                ("normal_coeff", ctypes.c_double),
                ("light", (ctypes.c_double * 8) * PP_MAX_LIGHTS),
                ("proj_hmoy", ctypes.c_double), ("proj_k", ctypes.c_double * 2)]


class Fieldlines_pp:
    """ Parameters of the reference's `Fieldlines_pp` (postproc.py:409-444): same
    constructor; `k_arr` / `phi_arr` as computed in its `__getitem__` (:455-462). """
    def __init__(self, n_iter=5, swirl=0., endpoint_k=1.0):
        if not 1 <= int(n_iter) <= PP_MAX_FL:
            raise ValueError(f"n_iter must be in 1..{PP_MAX_FL}")
        self.n_iter, self.swirl, self.endpoint_k = int(n_iter), float(swirl), float(endpoint_k)

    def arrays(self):
        k_arr = np.geomspace(1., self.endpoint_k, num=self.n_iter)
        k_arr = k_arr / np.sum(k_arr)
        rg = np.random.default_rng(0)
        phi_arr = rg.random(self.n_iter) * self.swirl * np.pi
        return k_arr, phi_arr


class Blinn_lighting:
    """ Scene light sources, constructor and `add_light_source` as the
    reference's (colors/layers.py:799-855).  `coefficients` are what crosses the
    C ABI: direction of each light and of its half-way vector (:866-888). """
    def __init__(self, k_ambient, color_ambient, **light_sources):
        self.k_ambient = k_ambient
        self.color_ambient = np.asarray(color_ambient)
        self.light_sources = []
        for ls in light_sources.values():
            self.add_light_source(**ls)

    def add_light_source(self, k_diffuse, k_specular, shininess, polar_angle, azimuth_angle,
                         color=np.array([1., 1., 1.]), material_specular_color=None):
        if len(self.light_sources) >= PP_MAX_LIGHTS:
            raise ValueError(f"at most {PP_MAX_LIGHTS} light sources")
        self.light_sources += [{
            "k_diffuse": np.asarray(k_diffuse), "k_specular": np.asarray(k_specular),
            "shininess": shininess, "polar_angle": polar_angle, "azimuth_angle": azimuth_angle,
            "color": np.asarray(color), "material_specular_color": material_specular_color}]

    def coefficients(self):
        out = []
        for ls in self.light_sources:
            theta = ls["polar_angle"] * np.pi / 180.
            phi = ls["azimuth_angle"] * np.pi / 180.
            phi_half = (np.pi * 0.5 + phi) * 0.5
            out.append([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi),
                        np.cos(theta) * np.cos(phi_half), np.sin(theta) * np.cos(phi_half),
                        np.sin(phi_half), float(ls["shininess"]),
                        float(np.any(np.asarray(ls["k_specular"]) != 0.))])
        return out

    def shade_XYZ(self, XYZ, shade):
        """ The colour arithmetic of `shade` / `partial_shade` (:857-903) on an XYZ
        image (ny, nx, 3) from the (2 n_lights, ny, nx) coefficient planes the GPU
        produced; the rgb <-> XYZ conversions stay with the caller's colour layers. """
        res = XYZ * self.k_ambient * self.color_ambient
        for l, ls in enumerate(self.light_sources):
            lambert, specular = shade[2 * l][:, :, np.newaxis], shade[2 * l + 1][:, :, np.newaxis]
            sp = XYZ if ls["material_specular_color"] is None else np.asarray(ls["material_specular_color"])
            res = res + (ls["k_diffuse"] * lambert * XYZ + ls["k_specular"] * specular * sp) * ls["color"]
        return res


def make_ext(fractal, calc_name, d, fieldlines=None, lighting=None, max_slope=70.):
    """ fsb_postproc_ext of one calculation: `fieldlines` a Fieldlines_pp,
    `lighting` a Blinn_lighting (+ the Normal_map_layer's max_slope, degrees). """
    x = FsbPostprocExt()
    state = fractal._calc_data[calc_name]["state"]
    codes = list(state.codes[0])
    if fieldlines is not None:
        k_arr, phi_arr = fieldlines.arrays()
        x.fl_n_iter = fieldlines.n_iter
        for i in range(fieldlines.n_iter):
            x.fl_k[i], x.fl_phi[i] = float(k_arr[i]), float(phi_arr[i])
        key = "zn_orbit" if d.holomorphic else "xn_orbit"
        backshift = getattr(fractal, "backshift", None)
        if key in codes and backshift is not None:
            x.fl_row_orbit, x.fl_backshift = codes.index(key), int(backshift)
        else:                      # postproc.py:470-481: start from zn, not backward
            x.fl_row_orbit, x.fl_backshift = -1, 0
        if d.holomorphic:
            x.fl_model = int(getattr(fractal, "exponent", 2))
        else:
            from .models import get_flavor_int
            x.fl_model = -int(get_flavor_int(fractal.flavor))
        pd = fractal.projection.c_abi_desc()       # proj_impl of get_std_cpt
        x.proj_kind, x.proj_hmoy = int(pd.kind), float(pd.hmoy)
        x.proj_k[0], x.proj_k[1] = float(pd.pix_to_ht[0]), float(pd.pix_to_ht[1])
        x.c_center[0], x.c_center[1] = float(fractal.x), float(fractal.y)
        x.c_scale = float(fractal.dx)
        for i, v in enumerate(np.asarray(fractal.lin_mat, np.float64).ravel()):
            x.c_lin_mat[i] = float(v)
    if lighting is not None:
        coeffs = lighting.coefficients()
        if not coeffs:
            raise ValueError("the lighting has no light source")
        x.n_lights = len(coeffs)
        x.normal_coeff = float(np.sin(max_slope * np.pi / 180.))
        for l, row in enumerate(coeffs):
            for i, v in enumerate(row):
                x.light[l][i] = float(v)
    return x


def _declare(lib):
    if getattr(lib, "_pp_declared", False):
        return lib
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    D = ctypes.POINTER(FsbPostprocDesc)
    X = ctypes.POINTER(FsbPostprocExt)
    lib.fsb_postproc_ext_run.argtypes = [D, X, i64, i32, vp, vp, vp, vp, vp]
    lib.fsb_postproc_ext_run_device.argtypes = [D, X, i64, i32, vp, vp, vp, vp, vp]
    lib.fsb_frame_run_grid_pp_ext.argtypes = [vp, i32, vp, vp, vp, D, X, vp, vp, vp, vp, vp, vp,
                                              vp, vp, vp, ctypes.POINTER(_native.FsbStats)]
    lib.fsb_frame_run_pp.argtypes = [vp, i32, vp, vp, i64, vp, D, vp, vp, vp, vp, vp, vp, vp,
                                     ctypes.POINTER(_native.FsbStats)]
    lib.fsb_frame_run_grid_pp.argtypes = [vp, i32, vp, vp, vp, D, vp, vp, vp, vp, vp, vp, vp,
                                          ctypes.POINTER(_native.FsbStats)]
    lib.fsb_postproc_run.argtypes = [D, i64, i32, vp, vp, vp, vp, vp, vp]
    lib.fsb_postproc_run_proj.argtypes = [D, i64, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.fsb_postproc_run_proj_device.argtypes = [D, i64, i32, vp, vp, vp, vp, vp, vp, vp]
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
    d.df_kind, k = projection_df(fractal.projection)
    d.df_k[0], d.df_k[1] = k.real, k.imag
    return d, dtype, len(codes)


def projection_df(proj):
    """ (df_kind, pix_to_ht) of fsb_postproc_desc: which derivative the
    reference's Postproc.get_dzndc applies for this projection
    (projection.py:188-192 Cartesian: none; :375-453 Expmap, stepped or not) """
    from . import projection as _projection
    if isinstance(proj, _projection.Expmap):
        k = complex(proj.pix_to_ht)
        if proj.use_step:
            return (1 if proj.rotates_df else 0), k
        return (2 if proj.rotates_df else 3), k
    return 0, 0j


def _check_projection(fractal, names):
    """ Projections cross the C ABI as parameters: anything else has no GPU form """
    from . import projection as _projection
    if not isinstance(fractal.projection, (_projection.Cartesian, _projection.Expmap)):
        raise NotImplementedError(f"no GPU post-processing for {type(fractal.projection).__name__}")


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


def _check_ext(fractal, d, fieldlines, lighting):
    if fieldlines is None and lighting is None:
        return
    _check_projection(fractal, ())
    if lighting is not None and d.row_dzndc < 0:
        raise ValueError("shading needs the derivative fields (calc_dzndc / calc_hessian)")


def fields_from_raw(fractal, calc_name, Z, stop_iter, fields=("cont_iter", "DEM", "normal"),
                    floor_iter=0, px_snap=None, dtype=np.float32, c_pix=None, fieldlines=None,
                    lighting=None, max_slope=70.):
    """ Stand-alone: raw (n_fields, npts) arrays on the host -> dict of fields.
    `fieldlines` (a Fieldlines_pp; needs `c_pix`, the pixel offsets of the points)
    adds "fieldlines"; `lighting` (a Blinn_lighting) adds "shade", the
    (2 n_lights, npts) Lambert / specular coefficients. """
    lib = _declare(_native.cuda_lib())
    d, dtype, n_rows = make_desc(fractal, calc_name, floor_iter, px_snap, dtype)
    Z = np.ascontiguousarray(Z)
    si = np.ascontiguousarray(np.ravel(stop_iter), dtype=np.int32)
    npts = Z.shape[1]
    out = _outputs(fields, npts, dtype, d.row_dzndc >= 0)
    _check_projection(fractal, out)
    _check_ext(fractal, d, fieldlines, lighting)
    cp = None
    if c_pix is not None:
        cp = np.ascontiguousarray(np.ravel(c_pix), dtype=np.complex128)
        if cp.shape[0] != npts:
            raise ValueError("c_pix does not match Z")
    if d.df_kind != 0 and d.row_dzndc >= 0 and cp is None \
            and ((set(out) & {"DEM", "normal_x", "normal_y"}) or lighting is not None):
        raise ValueError("this projection's derivative needs c_pix (the pixel offsets of the points)")
    if out:
        rc = lib.fsb_postproc_run_proj(ctypes.byref(d), npts, Z.shape[0], _native.ptr(Z),
                                       _native.ptr(si), _native.ptr(cp),
                                       *[_native.ptr(out.get(k)) for k in FIELDS])
        _native.check(lib, rc)
    if fieldlines is not None or lighting is not None:
        x = make_ext(fractal, calc_name, d, fieldlines, lighting, max_slope)
        if fieldlines is not None:
            if cp is None:
                raise ValueError("field lines need c_pix (the pixel offsets of the points)")
            out["fieldlines"] = np.empty(npts, dtype)
        if lighting is not None:
            out["shade"] = np.empty((2 * x.n_lights, npts), dtype)
        rc = lib.fsb_postproc_ext_run(ctypes.byref(d), ctypes.byref(x), npts, Z.shape[0],
                                      _native.ptr(Z), _native.ptr(si), _native.ptr(cp),
                                      _native.ptr(out.get("fieldlines")),
                                      _native.ptr(out.get("shade")))
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
    for spec in names:
        name, dt = spec[0], spec[1]
        shape = (st["npts"],) if len(spec) == 2 else (spec[2], st["npts"])
        cur = st["bufs"].get(name)
        if cur is None or cur.dtype != np.dtype(dt) or cur.shape != shape:
            if cur is not None:
                _native.pinned_free(cur)
            st["bufs"][name] = _native.pinned_empty(shape, dt)
    return st


def settings_chunk():
    from . import settings
    return settings.chunk_size


def frame_fields(fractal, calc_name, fields=("cont_iter", "DEM", "normal"), floor_iter=0,
                 px_snap=None, dtype=np.float32, want_stop_iter=False, copy=True, tiles=None,
                 fieldlines=None, lighting=None, max_slope=70.):
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
    _check_ext(fractal, d, fieldlines, lighting)
    want = [(k, dtype) for k in names] + [("stop_reason", np.int8)]
    if want_stop_iter:
        want.append(("stop_iter", np.int32))
    x = None
    if fieldlines is not None or lighting is not None:
        x = make_ext(fractal, calc_name, d, fieldlines, lighting, max_slope)
        if fieldlines is not None:
            want.append(("fieldlines", dtype))
        if lighting is not None:
            want.append(("shade", dtype, 2 * x.n_lights))
    st = _frame_staging(fractal, want, dtype, tiles)
    b, npts = st["bufs"], st["npts"]
    out = {w[0]: b[w[0]] for w in want}
    stats = _native.FsbStats()
    if x is None:
        rc = lib.fsb_frame_run_grid_pp(frame.ptr, st["tw"].shape[0], _native.ptr(st["tw"]),
                                       _native.ptr(st["th"]), _native.ptr(st["axes"]),
                                       ctypes.byref(d), *[_native.ptr(out.get(k)) for k in FIELDS],
                                       _native.ptr(out["stop_reason"]),
                                       _native.ptr(out.get("stop_iter")),
                                       _native.ptr(interrupted), stats)
    else:
        rc = lib.fsb_frame_run_grid_pp_ext(frame.ptr, st["tw"].shape[0], _native.ptr(st["tw"]),
                                           _native.ptr(st["th"]), _native.ptr(st["axes"]),
                                           ctypes.byref(d), ctypes.byref(x),
                                           *[_native.ptr(out.get(k)) for k in FIELDS],
                                           _native.ptr(out.get("fieldlines")),
                                           _native.ptr(out.get("shade")),
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
