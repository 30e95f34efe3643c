# -*- coding: utf-8 -*-
"""
`PerturbationFractal`: host side of the deep-zoom perturbation path.

Mirror of the marshalling part of the reference's
`fractalshades.perturbation.PerturbationFractal` (perturbation.py:22-977):
arbitrary-precision zoom parameters (mpmath), the reference orbit (native
MPFR, csrc/fp_orbit.c; cached in data/ref_pt.dat like the reference), the
per-frame scalars (`drift_xr`, `dx_xr`, `kc`, perturbation.py:166-194,317-375)
and the creation of the device frame that replaces the `cycle_indep_args`
tuple (perturbation.py:431-562).  dZndc / dZndz paths and the BLA tree are
computed by libfsb200 (fsb_frame_create); the pixel loop is
`numba_cycle_call` -> fsb_frame_run.
"""
import ctypes
import os
import pickle

import mpmath
import numpy as np

from . import settings
from . import xrange as fsx
from . import _native
from .core import Fractal, zoom_options

XR_CAP = 1 << 20


class FrameHandle:
    """ Owns one device frame (fsb_frame*) and the stats of its last run. """

    def __init__(self, lib, ptr, tables):
        self.lib = lib
        self.ptr = ptr
        self.tables = tables
        self.nz = lib.fsb_frame_nz(ptr)
        self.last_stats = None

    def close(self):
        if self.ptr:
            self.lib.fsb_frame_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup_ms(self):
        return {k: self.lib.fsb_frame_setup_ms(self.ptr, i)
                for i, k in enumerate(("upload", "dzndc_path", "bla_build"))}

    def get_bla(self):
        n = self.lib.fsb_frame_bla_len(self.ptr)
        width = 2 if self.tables["kind"] == "perturb_M2" else 8
        dt = np.complex128 if width == 2 else np.float64
        M = np.zeros(n * width, dt)
        r = np.zeros(n, np.float64)
        _native.check(self.lib, self.lib.fsb_frame_get_bla(
            self.ptr, _native.ptr(M), _native.ptr(r)))
        return M, r, n, self.lib.fsb_frame_stages_bla(self.ptr)

    def get_dzndc(self):
        L = len(self.tables["Zn_path"])
        if self.tables["kind"] == "perturb_M2":
            out = np.zeros(L, np.complex128)
            oe = np.zeros(L, np.int32)
        else:
            out = np.zeros((4, L), np.float64)
            oe = np.zeros((4, L), np.int32)
        _native.check(self.lib, self.lib.fsb_frame_get_dzndc(
            self.ptr, _native.ptr(out), _native.ptr(oe)))
        return out, (oe if self.tables["xr_detect"] else None)

    def get_dzndz(self):
        L = len(self.tables["Zn_path"])
        out = np.zeros(L + 1, np.complex128)
        oe = np.zeros(L + 1, np.int32)
        _native.check(self.lib, self.lib.fsb_frame_get_dzndz(
            self.ptr, _native.ptr(out), _native.ptr(oe)))
        return out, (oe if self.tables["xr_detect"] else None)

    def run(self, c_pix, Z, U, stop_reason, stop_iter, interrupted=None,
            tiles=None):
        from .core import TileAxes, check_outputs
        stats = _native.FsbStats()
        if isinstance(c_pix, TileAxes):
            # tile scheduler: pixel offsets expanded on the device from the axes
            check_outputs(c_pix.npts, Z, U, stop_reason, stop_iter, self.nz, None)
            rc = self.lib.fsb_frame_run_grid(
                self.ptr, c_pix.tw.shape[0], _native.ptr(c_pix.tw), _native.ptr(c_pix.th),
                _native.ptr(c_pix.axes), _native.ptr(Z), _native.ptr(U),
                _native.ptr(stop_reason), _native.ptr(stop_iter),
                _native.ptr(interrupted), stats)
            _native.check(self.lib, rc)
            self.last_stats = stats.as_dict()
            return rc
        npts = c_pix.shape[0]
        check_outputs(npts, Z, U, stop_reason, stop_iter, self.nz, c_pix)
        if tiles is not None:
            from .core import tile_shape_arrays
            tw, th = tile_shape_arrays(tiles, npts)
            rc = self.lib.fsb_frame_run_tiles(
                self.ptr, tw.shape[0], _native.ptr(tw), _native.ptr(th),
                _native.ptr(c_pix), _native.ptr(Z), _native.ptr(U),
                _native.ptr(stop_reason), _native.ptr(stop_iter),
                _native.ptr(interrupted), stats)
        else:
            rc = self.lib.fsb_frame_run(
                self.ptr, npts, _native.ptr(c_pix), _native.ptr(Z), _native.ptr(U),
                _native.ptr(stop_reason), _native.ptr(stop_iter),
                _native.ptr(interrupted), stats)
        _native.check(self.lib, rc)
        self.last_stats = stats.as_dict()
        return rc


def create_frame(tables, strict=None, use_tables=False):
    """ Build the device frame from a tables dict (see `frame_tables`).
    use_tables=True uploads the dict's own dZndc / dZndz / BLA arrays (staged
    parity against the oracle's tables); otherwise libfsb200 computes them. """
    lib = _native.cuda_lib(strict)
    t = tables
    d = _native.FsbFrameDesc()
    keep = []

    def arr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    m2 = (t["kind"] == "perturb_M2")
    d.model = _native.FSB_MODEL_M2 if m2 else _native.FSB_MODEL_BS
    d.flavor = int(t.get("flavor", 0))
    d.nexp = int(t.get("nexp", 0) or 0) if m2 else 0    # Perturbation_mandelbrot_N
    d.L = len(t["Zn_path"])
    d.Zn_path = arr(t["Zn_path"], np.complex128)
    idx = t.get("ref_index_xr")
    d.n_xr = 0 if idx is None else len(idx)
    d.ref_index_xr = arr(idx, np.int32)
    if m2:
        d.ref_xr = arr(t.get("ref_xr"), np.complex128)
        d.ref_xr_e = arr(t.get("ref_xr_e"), np.int32)
        dr = complex(t["drift"])
        d.drift[0], d.drift[1] = dr.real, dr.imag
        d.drift_e[0] = int(t["drift_e"])
        d.drift_e[1] = int(t["drift_e"])
        d.calc_dzndc = int(bool(t["calc_dzndc"]))
        d.calc_dzndz = int(bool(t["calc_dzndz"]))
        d.epsilon_stationnary_sq = float(t["epsilon_stationnary"]) ** 2
    else:
        d.ref_xr = arr(t.get("refx_xr"), np.float64)
        d.ref_xr_e = arr(t.get("refx_xr_e"), np.int32)
        d.refy_xr = arr(t.get("refy_xr"), np.float64)
        d.refy_xr_e = arr(t.get("refy_xr_e"), np.int32)
        d.drift[0], d.drift[1] = float(t["driftx"]), float(t["drifty"])
        d.drift_e[0], d.drift_e[1] = int(t["driftx_e"]), int(t["drifty_e"])
        d.calc_dzndc = int(bool(t["calc_hessian"]))
        d.calc_dzndz = 0
    d.ref_div_iter = int(t["ref_div_iter"])
    d.ref_order = int(t["ref_order"])
    d.lin_scale, d.lin_scale_e = float(t["lin_scale"]), int(t["lin_scale_e"])
    lm = np.asarray(t["lin_mat"], np.float64).ravel()
    for i in range(4):
        d.lin_mat[i] = lm[i]
    d.kc, d.kc_e = float(t["kc"]), int(t["kc_e"])
    # perturbation.py:459-468: dx_xr, times the projection's own scale if any
    d.scale_deriv = float(t.get("scale_deriv", t["dx"]))
    d.scale_deriv_e = int(t.get("scale_deriv_e", t["dx_e"]))
    pj = t.get("proj")
    if pj is not None:
        d.proj.kind = int(pj["kind"])
        d.proj.dzndc_modifier = int(pj["dzndc_modifier"])
        d.proj.hmoy = float(pj["hmoy"])
        d.proj.pix_to_ht[0], d.proj.pix_to_ht[1] = float(pj["k_re"]), float(pj["k_im"])
        d.proj.mod_param = float(pj["mod_param"])
    d.xr_detect = int(bool(t["xr_detect"]))
    d.bla_activated = int(bool(t["bla_activated"]))
    d.calc_orbit = int(bool(t.get("calc_orbit", False)))
    d.backshift = int(t.get("backshift", 0) or 0)
    d.max_iter = int(t["max_iter"])
    d.M_divergence_sq = float(t["M_divergence"]) ** 2
    d.BLA_eps = float(t["BLA_eps"]) if t.get("BLA_eps") is not None else 0.
    if use_tables:
        if m2:
            d.dZndc = arr(t.get("dZndc"), np.complex128)
            d.dZndc_e = arr(t.get("dZndc_e"), np.int32)
            d.dZndz = arr(t.get("dZndz"), np.complex128)
            d.dZndz_e = arr(t.get("dZndz_e"), np.int32)
        elif t.get("dXnda") is not None:
            d4 = np.stack([np.asarray(t[k], np.float64) for k in
                           ("dXnda", "dXndb", "dYnda", "dYndb")])
            d.dZndc = arr(d4, np.float64)
            if t.get("dXnda_e") is not None:
                e4 = np.stack([np.asarray(t[k + "_e"], np.int32) for k in
                               ("dXnda", "dXndb", "dYnda", "dYndb")])
                d.dZndc_e = arr(e4, np.int32)
        if t.get("M_bla") is not None:
            d.M_bla = arr(t["M_bla"], np.complex128 if m2 else np.float64)
            d.r_bla = arr(t["r_bla"], np.float64)
            d.bla_len = int(t["bla_len"])
            d.stages_bla = int(t["stages_bla"])
    out = ctypes.c_void_p()
    _native.check(lib, lib.fsb_frame_create(ctypes.byref(d), ctypes.byref(out)))
    return FrameHandle(lib, out, t)


def mpc_lin_proj_impl_noscale(lin_mat, x, y):
    """ perturbation.py:2682-2685 """
    x1 = lin_mat[0, 0] * x + lin_mat[0, 1] * y
    y1 = lin_mat[1, 0] * x + lin_mat[1, 1] * y
    return mpmath.mpc(x1, y1)


class PerturbationFractal(Fractal):

    def __init__(self, directory):
        super().__init__(directory)

    @zoom_options
    def zoom(self, *, precision: int, x, y, dx, nx: int, xy_ratio: float,
             theta_deg: float, projection=None, has_skew: bool = False,
             skew_00: float = 1., skew_01: float = 0., skew_10: float = 0.,
             skew_11: float = 1.):
        """ perturbation.py:43-135 """
        mpmath.mp.dps = precision
        self.x = mpmath.mpf(x)
        self.y = mpmath.mpf(y)
        self.dx = dx = mpmath.mpf(dx)
        self._set_projection(projection)
        self._skew = None
        if has_skew:
            self._skew = np.array(((skew_00, skew_01), (skew_10, skew_11)),
                                  dtype=np.float64)
        self.dx_std = float(dx)
        self.dx_xr = fsx.mpf_to_xr(dx)
        self.lin_scale_xr = self.dx_xr
        self.lin_mat = self.get_lin_mat()
        self.projection.adjust_to_zoom(self)
        pix = self.projection.min_local_scale * self.dx / self.nx
        with mpmath.workdps(6):
            required_dps = int(-mpmath.log10(pix / nx) + 1)
        if required_dps > precision:
            raise ValueError(
                "Precision is too low for min. pixel size and shall be "
                f"increased to {required_dps} (current setting: {precision}).")

    @property
    def xr_detect_activated(self):
        """ perturbation.py:152-155 """
        return bool(self.dx < settings.xrange_zoom_level)

    def ref_point_file(self):
        # `ref_point_dir`: orbit cache shared by fractals that live in different
        # directories (the frames of a zoom sequence)
        base = getattr(self, "ref_point_dir", None) or self.directory
        return os.path.join(base, "data", "ref_pt.dat")

    def ref_point_kc(self):
        """ perturbation.py:166-194 : bound on |dc| over the image, x 1.1 """
        w, h = self.projection.bounding_box(self.xy_ratio)
        dx = self.dx
        mat = self.lin_mat
        corners = [mpc_lin_proj_impl_noscale(mat, sx * 0.5 * w, sy * 0.5 * h) * dx
                   for sx, sy in ((1, 1), (-1, 1), (-1, -1), (1, -1))]
        c0 = self.x + 1j * self.y
        shift = self.FP_params["ref_point"] - c0
        kc = max(abs(shift - c) for c in corners) * 1.1
        return fsx.mpf_to_xr(kc)

    # -- reference orbit -----------------------------------------------------
    def ref_point_matching(self):
        """ perturbation.py:211-253 """
        init_kwargs = self.init_kwargs
        del init_kwargs["directory"]
        try:
            FP = self.FP_params
        except (FileNotFoundError, EOFError):
            return False
        drift = (self.x + 1j * self.y) - FP["ref_point"]
        with mpmath.workdps(30):
            loc = abs(drift / self.dx) ** 2 < 1.e6
        match = (mpmath.mp.dps <= FP["dps"] + 3 and bool(loc)
                 and FP["max_iter"] >= self.max_iter
                 and all(init_kwargs.get(k) == v
                         for k, v in FP["init_kwargs"].items())
                 # an orbit registered with its Xrange points serves shallower
                 # frames too (zoom movies); the converse does not hold
                 and (bool(FP.get("xr_detect", False)) or not self.xr_detect_activated))
        return bool(match)

    def save_ref_point(self, FP_params, Zn_path):
        self._FP_params = FP_params
        self._Zn_path = Zn_path
        path = self.ref_point_file()
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp"      # atomic: other ranks may be reading
        with open(tmp, 'wb+') as tmpfile:
            pickle.dump(FP_params, tmpfile, pickle.HIGHEST_PROTOCOL)
            pickle.dump(Zn_path, tmpfile, pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, path)

    def reload_ref_point(self, scan_only=False):
        with open(self.ref_point_file(), 'rb') as tmpfile:
            FP_params = pickle.load(tmpfile)
            if scan_only:
                return FP_params
            Zn_path = pickle.load(tmpfile)
        return FP_params, Zn_path

    @property
    def FP_params(self):
        if not hasattr(self, "_FP_params"):
            self._FP_params = self.reload_ref_point(scan_only=True)
        return self._FP_params

    @property
    def Zn_path(self):
        if not hasattr(self, "_Zn_path"):
            self._FP_params, self._Zn_path = self.reload_ref_point()
        return self._Zn_path

    def get_FP_orbit(self, c0=None, newton="cv", order=None, max_newton=None):
        """ perturbation.py:651-768 : reference point = the nucleus found by
        the ball method + Newton descent around the image centre (periodic
        reference, `ref_order` wrap), or the image centre itself when
        settings.no_newton is set or the descent fails.  Every model has its
        native search (holomorphic power 2 / power N, burning-ship family). """
        if newton == "step":
            raise NotImplementedError("step option not Implemented (yet)")
        if self.ref_point_matching():
            return
        if self.dx > settings.newton_zoom_level:
            self.compute_critical_orbit(self.critical_pt)
            return
        if c0 is None:
            c0 = self.x + 1j * self.y
        if settings.no_newton or (newton is None) or (newton == "None"):
            self.compute_FP_orbit(c0, None)
            return
        if order is None:
            order = self.ball_method(c0, self.dx * 1.0)
            if order is None:            # ball method failed: image centre
                self.compute_FP_orbit(c0, None)
                return
        max_attempt = 2
        eps_pixel = self.dx * (1. / self.nx)

        def descent(no_div_allowed=True):
            if no_div_allowed:
                try:
                    return True, self.find_nucleus(c0, order, eps_pixel,
                                                   max_newton=max_newton)
                except NotImplementedError:
                    pass
            return False, self.find_any_nucleus(c0, order, eps_pixel,
                                                max_newton=max_newton)
        _, (newton_cv, nucleus) = descent()
        attempt = 1
        if not newton_cv:
            while (not newton_cv) and attempt <= max_attempt:
                attempt += 1
                mpmath.mp.dps = int(1.25 * mpmath.mp.dps)
                eps_pixel = self.dx * (1. / self.nx)
                no_div, (newton_cv, nucleus) = descent()
                if no_div and (not newton_cv) and (attempt == max_attempt):
                    # last try: accept the cycles of the divisors of the order
                    newton_cv, nucleus = self.find_any_nucleus(
                        c0, order, eps_pixel, max_newton=max_newton)
        if not newton_cv:
            order = None                 # the reference cannot be wrapped
            nucleus = c0
        self.compute_FP_orbit(nucleus, order)

    def ball_method(self, c, px, kind=1, M_divergence=1.e5):
        """ perturbation.py:857-867 : first period of the nucleus around c """
        if kind != 1:
            raise NotImplementedError("ball method kind 2")
        return self._ball_method(c, px, self.max_iter, M_divergence)

    def _fp_base(self, ref_point, order):
        init_kwargs = self.init_kwargs
        del init_kwargs["directory"]
        return {"ref_point": ref_point, "dps": mpmath.mp.dps, "order": order,
                "max_iter": self.max_iter, "FP_code": self.FP_code,
                "init_kwargs": init_kwargs,
                "xr_detect": self.xr_detect_activated}

    def compute_critical_orbit(self, crit):
        """ perturbation.py:770-805 : shallow zoom, 101 copies of the critical
        point, order 100 """
        FP = self._fp_base(crit, 100)
        FP["partials"] = {}
        FP["xr"] = {}
        FP["div_iter"] = 100
        Zn_path = np.zeros([101], dtype=np.complex128)
        Zn_path[:] = crit
        self.save_ref_point(FP, Zn_path)

    def compute_FP_orbit(self, ref_point, order=None):
        """ perturbation.py:808-854 """
        max_iter = self.max_iter
        FP = self._fp_base(ref_point, order)
        ref_orbit_len = max_iter + 1
        if order is not None:
            ref_orbit_len = min(order, ref_orbit_len)
        FP["ref_orbit_len"] = ref_orbit_len
        # (the reference uses np.empty: entries past the escape index are never
        # read; zeros keep the BLA table deterministic there)
        Zn_path = np.zeros([ref_orbit_len], dtype=np.complex128)
        i, partial_dict, xr_dict = self.FP_loop(Zn_path, ref_point)
        FP["partials"] = partial_dict
        FP["xr"] = xr_dict
        FP["div_iter"] = i
        self.save_ref_point(FP, Zn_path)

    def _native_orbit(self, NP_orbit, c0, flavor=None, exponent=2):
        """ Native MPFR orbit (include/fsb200_orbit.h); returns the reference's
        3-tuple (div_iter, partials{}, xr{}) with xr values as raw
        (mx, ex, my, ey) parts. """
        lib = _native.load_orbit_lib()
        orbit = NP_orbit.view(np.float64)
        max_orbit_iter = NP_orbit.shape[0] - 1
        buf = (_native.OrbitXr * XR_CAP)()
        cnt = ctypes.c_int64(0)
        seed_prec = mpmath.mp.prec
        sx = str(c0.real).encode('utf8')
        sy = str(c0.imag).encode('utf8')
        M = self.M_divergence * 2     # models/mandelbrot_M2.py:425
        if flavor is None:
            i = lib.fsb_orbit_mandelbrot(
                orbit.ctypes.data, max_orbit_iter, exponent,
                int(self.xr_detect_activated), M, sx, sy, seed_prec, buf,
                XR_CAP, ctypes.byref(cnt))
        else:
            i = lib.fsb_orbit_burning_ship(
                orbit.ctypes.data, max_orbit_iter, flavor,
                int(self.xr_detect_activated), M, sx, sy, seed_prec, buf,
                XR_CAP, ctypes.byref(cnt))
        if i < 0:
            raise RuntimeError(f"native reference orbit failed with code {i}")
        xr = {int(buf[k].index): (buf[k].mx, buf[k].ex, buf[k].my, buf[k].ey)
              for k in range(cnt.value)}
        return int(i), {}, xr

    # -- per-frame tables ------------------------------------------------------
    def get_path_data(self):
        """ perturbation.py:317-375 (plain dict instead of a tuple) """
        FP = self.FP_params
        Zn_path = self.Zn_path
        xr_py = FP["xr"]
        ref_order = FP["order"]
        ref_div_iter = FP["div_iter"]
        if ref_order is not None:
            ref_div_iter = self.max_iter + 1
        t = {"Zn_path": Zn_path, "ref_div_iter": int(ref_div_iter),
             "ref_order": int(ref_order) if ref_order is not None else (1 << 62)}
        t["dx"], t["dx_e"] = fsx.mpf_to_xr(self.dx)
        n = len(xr_py)
        idx = np.array(sorted(xr_py.keys()), dtype=np.int32)
        t["ref_index_xr"] = idx if n > 0 else None
        if self.holomorphic:
            t["drift"], t["drift_e"] = fsx.mpc_to_xr(
                (self.x + 1j * self.y) - FP["ref_point"])
            if n > 0:
                vals = [fsx.xr_complex_from_parts(*xr_py[int(k)]) for k in idx]
                t["ref_xr"] = np.array([v[0] for v in vals], np.complex128)
                t["ref_xr_e"] = np.array([v[1] for v in vals], np.int32)
            else:
                t["ref_xr"] = t["ref_xr_e"] = None
        else:
            ref = FP["ref_point"]
            t["driftx"], t["driftx_e"] = fsx.mpf_to_xr(self.x - mpmath.mpf(ref.real))
            t["drifty"], t["drifty_e"] = fsx.mpf_to_xr(self.y - mpmath.mpf(ref.imag))
            if n > 0:
                t["refx_xr"] = np.array([xr_py[int(k)][0] for k in idx], np.float64)
                t["refx_xr_e"] = np.array([xr_py[int(k)][1] for k in idx], np.int32)
                t["refy_xr"] = np.array([xr_py[int(k)][2] for k in idx], np.float64)
                t["refy_xr_e"] = np.array([xr_py[int(k)][3] for k in idx], np.int32)
            else:
                t["refx_xr"] = t["refx_xr_e"] = t["refy_xr"] = t["refy_xr_e"] = None
        return t

    def frame_tables(self):
        """ Everything fsb_frame_create needs, as a plain dict (host only: the
        orbit and the Xrange scalars; no GPU is touched).  The same dict feeds
        the CPU oracle in the test-suite. """
        self.get_FP_orbit()
        t = self.get_path_data()
        t.update(self._kernel_options)
        t["xr_detect"] = self.xr_detect_activated
        t["lin_mat"] = np.array(self.lin_mat, np.float64)
        t["lin_scale"], t["lin_scale_e"] = self.lin_scale_xr
        # perturbation.py:459-468 : scale of the derivatives = dx, times the
        # projection-induced scale (Xrange_array product: mantissa product,
        # exponents added, renormalised)
        t["scale_deriv"], t["scale_deriv_e"] = t["dx"], t["dx_e"]
        if self.projection.scale != 1.:
            sm, se = fsx.mpf_to_xr(self.projection.scale)
            m, k = np.frexp(t["dx"] * sm)
            t["scale_deriv"], t["scale_deriv_e"] = float(m), int(t["dx_e"] + se + k)
        pd = self.projection.c_abi_desc()
        t["proj"] = dict(kind=pd.kind, dzndc_modifier=pd.dzndc_modifier, hmoy=pd.hmoy,
                         k_re=pd.pix_to_ht[0], k_im=pd.pix_to_ht[1],
                         mod_param=pd.mod_param)
        self.kc = kc = self.ref_point_kc()
        if kc[0] == 0.:
            raise RuntimeError("Resolution is too low for this zoom depth.")
        t["kc"], t["kc_e"] = kc
        return t

    def get_cycle_indep_args(self, initialize, iterate):
        """ perturbation.py:431-562 : reference orbit, derivative paths, BLA
        tree -> one device frame """
        self._kernel_options = vars(iterate).copy()
        tables = self.frame_tables()
        self._frame_tables = tables
        frame = create_frame(tables)
        return ("perturb", frame, self._interrupted)

    def reset_bla_tree(self, cycle_indep_args):
        """ perturbation.py:565-580 : new frame after `projection.set_exp_zoom_step`
        (BLA validity radii and derivative scale of the step); like the
        reference it simply rebuilds the frame with the same options """
        self._release_indep_args(cycle_indep_args)
        tables = self.frame_tables()
        self._frame_tables = tables
        return ("perturb", create_frame(tables), self._interrupted)

    def _release_indep_args(self, indep):
        if indep is not None and indep[0] == "perturb":
            indep[1].close()

    @staticmethod
    def numba_cycle_call(cycle_dep_args, cycle_indep_args, tiles=None):
        """ perturbation.py:414-428 : per-tile entry point, in-place
        (`tiles`: see Fractal.numba_cycle_call) """
        (kind, frame, interrupted) = cycle_indep_args
        (c_pix, Z, U, stop_reason, stop_iter) = cycle_dep_args
        rc = frame.run(c_pix, Z, U, stop_reason, stop_iter, interrupted, tiles)
        Fractal._last_stats = frame.last_stats
        return rc
