# -*- coding: utf-8 -*-
"""
Tile runtime of the hot path: the `Fractal` base class.

Host-side mirror of the slice of the reference's `fractalshades.core.Fractal`
that surrounds the per-pixel kernels (core.py:1264-2791): zoom parameters,
200x200 tiling, pixel offsets, `calc_hook`, report / data memmaps in the
reference's on-disk layout, and the tile loop.  The five methods the GPU
scheduler replaces keep their names:

    numba_cycle_call      core.py:2022-2027   -> one C-ABI call (GPU kernels)
    get_cycle_indep_args  core.py:2029-2041   -> per-frame descriptor
    get_cycling_dep_args  core.py:2044-2075   -> caller-owned tile arrays
    compute_rawdata_dev   core.py:2515-2554   -> GPU tile scheduler
    evaluate_rawdata_final core.py:2570-2592  -> on-the-fly tile

The reference dispatches one numba call per tile from a thread pool
(mthreading.py:48-68); here `compute_rawdata_dev` batches all pending tiles of
the frame into one launch (tiles are consecutive 1-D slabs of the memmaps, so
a batch is a single contiguous range) and the persistent kernel load-balances
at warp granularity.
"""
import concurrent.futures
import inspect
import fnmatch
import functools
import os
import pickle
import threading
import types

import numpy as np
from numpy.lib.format import open_memmap

from . import settings
from . import projection as _projection
from . import _native

USER_INTERRUPTED = 1


# ---------------------------------------------------------------------------
# decorators, same contract as the reference's utils.py:247-339
def zoom_options(method):
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        if len(args) > 0:
            raise TypeError(f"{method.__name__} should only accept "
                            f"keyword-arguments ; given positionnal: {args}")
        ba = inspect.signature(method).bind_partial(self, **kwargs)
        ba.apply_defaults()
        kwargs_dic = dict(ba.arguments)
        kwargs_dic.pop("self")
        self.zoom_kwargs = kwargs_dic
        for key, val in kwargs_dic.items():
            setattr(self, key, val)
        return method(self, **kwargs)
    wrapper._is_zoom_options = True
    return wrapper


def calc_options(method):
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        if len(args) > 0:
            raise TypeError(f"{method.__name__} should only accept "
                            f"keyword-arguments ; given positionnal: {args}")
        ba = inspect.signature(method).bind_partial(self, **kwargs)
        ba.apply_defaults()
        kwargs_dic = dict(ba.arguments)
        kwargs_dic.pop("self")
        return_dic = method(self, **kwargs)
        self.calc_hook(method.__name__, kwargs_dic, return_dic)
    wrapper._is_calc_options = True
    return wrapper


class KernelSpec(types.SimpleNamespace):
    """ What `initialize()` / `iterate()` return instead of a numba closure:
    a plain description of the device kernel variant (model, flavour, flags).
    Arbitrary Python callables cannot cross the C ABI. """


def dic_flatten(dic, key_prefix="", sep="@"):
    res = {}
    for k, v in dic.items():
        key = key_prefix + sep + str(k) if key_prefix else str(k)
        if isinstance(v, dict):
            res.update(dic_flatten(v, key, sep))
        else:
            res[key] = v
    return res


_TLS = threading.local()


class _FractalMeta(type):
    """ `Fractal._last_stats`: counters of the calling thread's last seam call
    (the seam is thread-safe, one launch context per host thread; a class
    global would be overwritten by concurrent tile threads). """

    @property
    def _last_stats(cls):
        return getattr(_TLS, "stats", None)

    @_last_stats.setter
    def _last_stats(cls, value):
        _TLS.stats = value


class Fractal(metaclass=_FractalMeta):
    REPORT_ITEMS = ["chunk1d_begin", "chunk1d_end", "chunk_pts", "done"]
    SAVE_ARRS = ["Z", "U", "stop_reason", "stop_iter"]
    USER_INTERRUPTED = USER_INTERRUPTED

    def __init__(self, directory: str):
        self.directory = directory
        self.subset = None
        self._interrupted = np.array([0], dtype=np.bool_)   # core.py:1352
        self.float_postproc_type = np.dtype(settings.postproc_dtype)
        self.termination_type = np.int8
        self.int_type = np.int32
        self.last_stats = None

    # -- init kwargs (fingerprint) ----------------------------------------
    @property
    def init_kwargs(self):
        return {p: getattr(self, p) for p in
                inspect.signature(self.__init__).parameters}

    # -- zoom ---------------------------------------------------------------
    @zoom_options
    def zoom(self, *, x: float = 0., y: float = 0., dx: float = 8.,
             nx: int = 800, xy_ratio: float = 1., theta_deg: float = 0.,
             projection=None, has_skew: bool = False, skew_00: float = 1.0,
             skew_01: float = 0.0, skew_10: float = 0.0, skew_11: float = 1.0):
        """ core.py:1402-1467 """
        if isinstance(x, str) or isinstance(y, str) or isinstance(dx, str):
            raise RuntimeError("Float expected for x, y, dx")
        self._set_projection(projection)
        self._skew = None
        if has_skew:
            self._skew = np.array(((skew_00, skew_01), (skew_10, skew_11)),
                                  dtype=np.float64)
        self.lin_mat = self.get_lin_mat()
        self.projection.adjust_to_zoom(self)

    def _set_projection(self, projection):
        """ Cartesian and Expmap reduce to the parameters of fsb_proj_desc;
        anything else cannot cross the C ABI and is refused (no fallback) """
        if projection is None or isinstance(projection, str):
            projection = _projection.Cartesian()
        if not isinstance(projection, _projection.Projection):
            raise NotImplementedError(
                f"projection {type(projection).__name__} is not supported by "
                "the GPU path (Cartesian and Expmap cross the C ABI; no fallback)")
        if type(projection).c_abi_desc is _projection.Projection.c_abi_desc:
            projection.c_abi_desc()                 # raises: no parametric form
        self.projection = projection
        self.zoom_kwargs["projection"] = projection.fingerprint()

    def get_lin_mat(self):
        """ core.py:1470-1484 """
        theta = self.theta_deg / 180. * np.pi
        c = np.cos(theta)
        s = np.sin(theta)
        lin_mat = np.array(((c, -s), (s, c)), dtype=np.float64)
        if self._skew is not None:
            lin_mat = np.matmul(self._skew, lin_mat)
        return lin_mat

    @property
    def ny(self):
        return int(self.nx / self.xy_ratio + 0.5)      # core.py:1531-1533

    @property
    def dy(self):
        return self.dx / self.xy_ratio

    @property
    def skew(self):
        return getattr(self, "_skew", None)

    @property
    def float_type(self):
        return np.float64

    # -- tiling, core.py:1644-1698 -------------------------------------------
    def chunk_slices(self):
        cs = settings.chunk_size
        for ix in range(0, self.nx, cs):
            ixx = min(ix + cs, self.nx)
            for iy in range(0, self.ny, cs):
                iyy = min(iy + cs, self.ny)
                yield (ix, ixx, iy, iyy)

    @property
    def chunks_count(self):
        cs = settings.chunk_size
        cx = -(-self.nx // cs)
        cy = -(-self.ny // cs)
        return cx * cy

    def chunk_rank(self, chunk_slice):
        cs = settings.chunk_size
        (ix, _, iy, _) = chunk_slice
        cy = -(-self.ny // cs)
        return (ix // cs) * cy + (iy // cs)

    def chunk_from_rank(self, rank):
        cs = settings.chunk_size
        cy = -(-self.ny // cs)
        cix, ciy = divmod(rank, cy)
        ix, iy = cix * cs, ciy * cs
        return (ix, min(ix + cs, self.nx), iy, min(iy + cs, self.ny))

    def tile_axes(self, chunk_slice):
        """ The two axes of a tile's pixel grid without jitter / supersampling:
        `chunk_pixel_pos(cs, False, None)[r, c] == x[c] + 1j * y[r]`, bit for
        bit (the grid calls of libfsb200 expand them on the device). """
        data_type = self.float_type
        (nx, ny) = (self.nx, self.ny)
        (ix, ixx, iy, iyy) = chunk_slice
        kx = 0.5 / (nx - 1)
        ky = 0.5 / (ny - 1)
        x_1d = np.linspace(kx * (2 * ix - nx + 1), kx * (2 * ixx - nx - 1),
                           num=(ixx - ix), dtype=data_type)
        y_1d = np.linspace(ky * (2 * iy - ny + 1), ky * (2 * iyy - ny - 1),
                           num=(iyy - iy), dtype=data_type)
        y = -y_1d
        y /= self.xy_ratio
        return x_1d, y

    def chunk_pixel_pos(self, chunk_slice, jitter, supersampling):
        """ core.py:1767-1830 : pixel offsets in fractions of dx, row 0 = top """
        data_type = self.float_type
        (nx, ny) = (self.nx, self.ny)
        (ix, ixx, iy, iyy) = chunk_slice
        kx = 0.5 / (nx - 1)
        ky = 0.5 / (ny - 1)
        if supersampling is None:
            x_1d = np.linspace(kx * (2 * ix - nx + 1), kx * (2 * ixx - nx - 1),
                               num=(ixx - ix), dtype=data_type)
            y_1d = np.linspace(ky * (2 * iy - ny + 1), ky * (2 * iyy - ny - 1),
                               num=(iyy - iy), dtype=data_type)
        else:
            ssg = supersampling
            ssg_gap = (ssg - 1.) / ssg
            x_1d = np.linspace(kx * (2 * ix - nx + 1 - ssg_gap),
                               kx * (2 * ixx - nx - 1 + ssg_gap),
                               num=(ixx - ix) * ssg, dtype=data_type)
            y_1d = np.linspace(ky * (2 * iy - ny + 1 - ssg_gap),
                               ky * (2 * iyy - ny - 1 + ssg_gap),
                               num=(iyy - iy) * ssg, dtype=data_type)
        dx_screen, dy_screen = np.meshgrid(x_1d, -y_1d, indexing='xy')
        if jitter:
            rg = np.random.default_rng(0)
            rand_x = rg.random(dx_screen.shape, dtype=data_type)
            rand_y = rg.random(dy_screen.shape, dtype=data_type)
            k = 0.7071067811865476
            jitter_x = (0.5 - rand_x) * k / (nx - 1) * jitter
            jitter_y = (0.5 - rand_y) * k / (ny - 1) * jitter
            if supersampling is not None:
                jitter_x /= supersampling
                jitter_y /= supersampling
            dx_screen += jitter_x
            dy_screen += jitter_y
        dy_screen /= self.xy_ratio
        return dx_screen + 1j * dy_screen

    def pts_count(self, calc_name, chunk_slice=None):
        state = self._calc_data[calc_name]["state"]
        subset = state.subset
        if subset is not None:
            if chunk_slice is None:
                return sum(int(np.count_nonzero(subset[cs]))
                           for cs in self.chunk_slices())
            return int(np.count_nonzero(subset[chunk_slice]))
        if chunk_slice is None:
            return self.nx * self.ny
        (ix, ixx, iy, iyy) = chunk_slice
        return (ixx - ix) * (iyy - iy)

    # -- calc hook, core.py:1918-2009 ----------------------------------------
    def calc_hook(self, calc_callable, calc_kwargs, return_dic):
        calc_name = calc_kwargs["calc_name"]
        if not hasattr(self, "_calc_data"):
            self._calc_data = dict()
        state = types.SimpleNamespace()
        for k, v in calc_kwargs.items():
            setattr(state, k, v)
            setattr(self, k, v)
        set_state = return_dic["set_state"]()
        set_state(state)
        set_state(self)
        initialize = return_dic["initialize"]()
        iterate = return_dic["iterate"]()
        old = self._calc_data.get(calc_name)
        if old is not None:
            self._release_indep_args(old.get("cycle_indep_args"))
        cycle_indep_args = self.get_cycle_indep_args(initialize, iterate)
        self._calc_data[calc_name] = {
            "calc_class": type(self).__name__,
            "calc_callable": calc_callable,
            "calc_kwargs": calc_kwargs,
            "zoom_kwargs": self.zoom_kwargs,
            "state": state,
            "cycle_indep_args": cycle_indep_args,
            "saved_codes": self.saved_codes(state.codes),
            "init_kwargs": self.init_kwargs,
        }
        fp_items = ("calc_class", "calc_callable", "calc_kwargs", "zoom_kwargs",
                    "init_kwargs")
        state.fingerprint = {k: self._calc_data[calc_name][k] for k in fp_items}
        if self.res_available(calc_name):
            try:
                for key in ["report"] + self.SAVE_ARRS:
                    path = (self.report_path(calc_name) if key == "report"
                            else self.data_path(calc_name)[key])
                    open_memmap(filename=path, mode="r+")
                self._calc_data[calc_name]["need_new_mmap"] = False
            except FileNotFoundError:
                self._calc_data[calc_name]["need_new_mmap"] = True
        else:
            self._calc_data[calc_name]["need_new_mmap"] = True
            self.save_fingerprint(calc_name, state.fingerprint)

    def _release_indep_args(self, indep):
        pass

    def raise_interruption(self):
        self._interrupted[0] = True

    def lower_interruption(self):
        self._interrupted[0] = False

    def is_interrupted(self):
        return bool(self._interrupted[0] or settings.skip_calc)

    # -- the seam -------------------------------------------------------------
    @staticmethod
    def numba_cycle_call(cycle_dep_args, cycle_indep_args, tiles=None):
        """ core.py:2022-2027.  Same in-place semantics: the caller-owned
        arrays of `cycle_dep_args` are filled; returns 0 or USER_INTERRUPTED.
        The work is done by libfsb200 (fsb_std_run), never on the CPU.

        tiles (optional, not in the reference): [(width, height), ...] when
        the point list is a concatenation of full row-major tiles; same
        results, better lane occupancy on the GPU (fsb_std_run_tiles). """
        (c_pix, Z, U, stop_reason, stop_iter) = cycle_dep_args
        (kind, desc, interrupted) = cycle_indep_args
        assert kind == "std"
        lib = _native.cuda_lib()
        stats = _native.FsbStats()
        if isinstance(c_pix, TileAxes):
            # tile scheduler: pixel offsets expanded on the device from the axes
            check_outputs(c_pix.npts, Z, None, stop_reason, stop_iter,
                          lib.fsb_std_nz(desc), None)
            rc = lib.fsb_std_run_grid(
                desc, c_pix.tw.shape[0], _native.ptr(c_pix.tw), _native.ptr(c_pix.th),
                _native.ptr(c_pix.axes), _native.ptr(Z), _native.ptr(stop_reason),
                _native.ptr(stop_iter), _native.ptr(interrupted), stats)
            _native.check(lib, rc)
            Fractal._last_stats = stats.as_dict()
            return rc
        npts = c_pix.shape[0]
        check_outputs(npts, Z, None, stop_reason, stop_iter, lib.fsb_std_nz(desc), c_pix)
        if tiles is not None:
            tw, th = tile_shape_arrays(tiles, npts)
            rc = lib.fsb_std_run_tiles(
                desc, tw.shape[0], _native.ptr(tw), _native.ptr(th),
                _native.ptr(c_pix), _native.ptr(Z), _native.ptr(stop_reason),
                _native.ptr(stop_iter), _native.ptr(interrupted), stats)
        else:
            rc = lib.fsb_std_run(desc, npts, _native.ptr(c_pix), _native.ptr(Z),
                                 _native.ptr(stop_reason), _native.ptr(stop_iter),
                                 _native.ptr(interrupted), stats)
        _native.check(lib, rc)
        Fractal._last_stats = stats.as_dict()
        return rc

    def get_cycle_indep_args(self, initialize, iterate):
        """ core.py:2029-2041 : digest of the zoom and calculation parameters """
        spec = iterate
        d = _native.FsbStdDesc()
        d.model = spec.model
        d.flavor = getattr(spec, "flavor", 0)
        d.center_re = float(self.x)
        d.center_im = float(self.y)
        d.dx = float(self.dx)
        lm = np.asarray(self.lin_mat, np.float64).ravel()
        for i in range(4):
            d.lin_mat[i] = lm[i]
        d.max_iter = int(spec.max_iter)
        d.M_divergence_sq = float(spec.M_divergence) ** 2
        d.epsilon_stationnary_sq = float(getattr(spec, "epsilon_stationnary", 0.)) ** 2
        d.calc_d2zndc2 = int(bool(getattr(spec, "calc_d2zndc2", False)))
        d.calc_orbit = int(bool(spec.calc_orbit))
        d.backshift = int(spec.backshift or 0)
        d.nexp = int(getattr(spec, "nexp", 0) or 0)         # Mandelbrot_N
        # pixel projection; the standard loops apply no dz/dc modifier
        # (core.py:2035: the reference's own hook is commented out)
        d.proj = self.projection.c_abi_desc()
        d.proj.dzndc_modifier = _native.FSB_DZNDC_MOD_NONE
        d.proj.mod_param = 0.
        return ("std", d, self._interrupted)

    def get_cycling_dep_args(self, calc_name, chunk_slice, final=False,
                             jitter=False, supersampling=None):
        """ core.py:2044-2075 """
        c_pix = np.ravel(self.chunk_pixel_pos(chunk_slice, jitter, supersampling))
        state = self._calc_data[calc_name]["state"]
        subset = state.subset
        if subset is not None:
            chunk_subset = np.asarray(subset[chunk_slice], dtype=bool)
            c_pix = c_pix[chunk_subset]
        else:
            chunk_subset = None
        c_pix = np.ascontiguousarray(c_pix)
        (n_pts,) = c_pix.shape
        n_Z, n_U, n_stop = (len(code) for code in state.codes)
        Z = np.zeros([n_Z, n_pts], dtype=state.complex_type)
        U = np.zeros([n_U, n_pts], dtype=self.int_type)
        stop_reason = - np.ones([1, n_pts], dtype=self.termination_type)
        stop_iter = np.zeros([1, n_pts], dtype=self.int_type)
        return (c_pix, Z, U, stop_reason, stop_iter), chunk_subset

    # -- fingerprint / files ---------------------------------------------------
    def fingerprint_path(self, calc_name):
        return os.path.join(self.directory, "data", calc_name + ".fingerprint")

    def save_fingerprint(self, calc_name, fingerprint):
        path = self.fingerprint_path(calc_name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        # several ranks may share the directory (tile sharding): the file
        # appears atomically, a reader never sees it half written
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, 'wb+') as fp_file:
            pickle.dump(_picklable(fingerprint), fp_file, pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, path)

    def reload_fingerprint(self, calc_name):
        with open(self.fingerprint_path(calc_name), 'rb') as tmpfile:
            return pickle.load(tmpfile)

    def fingerprint_matching(self, calc_name, test_fingerprint, log=False):
        flatten_fp = dic_flatten(test_fingerprint)
        state = self._calc_data[calc_name]["state"]
        expected_fp = dic_flatten(_picklable(state.fingerprint))
        for key, val in expected_fp.items():
            if key not in flatten_fp or flatten_fp[key] != val:
                return False
        return True

    def res_available(self, calc_name, chunk_slice=None):
        """ core.py:1883-1915 """
        try:
            fingerprint = self.reload_fingerprint(calc_name)
        except IOError:
            return False
        if not self.fingerprint_matching(calc_name, fingerprint):
            return False
        if chunk_slice is None:
            return True
        try:
            mmap = self.get_report_memmap(calc_name, mode="r")
        except IOError:
            return False
        return mmap[self.chunk_rank(chunk_slice), self.REPORT_ITEMS.index("done")] > 0

    def saved_codes(self, codes):
        (complex_codes, int_codes, stop_codes) = codes
        return (self.filter_stored_codes(complex_codes),
                self.filter_stored_codes(int_codes), stop_codes)

    @staticmethod
    def filter_stored_codes(codes):
        return [c for c in codes if not c.startswith("_")]

    def report_path(self, calc_name):
        return os.path.join(self.directory, "data", calc_name + ".report")

    def data_path(self, calc_name):
        keys = ["subset"] + self.SAVE_ARRS
        return {k: os.path.join(self.directory, "data", f"{calc_name}_{k}.arr")
                for k in keys}

    def init_report_mmap(self, calc_name):
        """ core.py:2174-2214 """
        items = self.REPORT_ITEMS
        n = self.chunks_count
        path = self.report_path(calc_name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        mmap = open_memmap(filename=path, mode='w+', dtype=np.dtype(np.int32),
                           shape=(n, len(items)), fortran_order=False)
        mmap[:, items.index("done")] = 0
        chunk_pts = np.empty((n,), dtype=np.int32)
        for i, chunk_slice in enumerate(self.chunk_slices()):
            chunk_pts[i] = self.pts_count(calc_name, chunk_slice)
        mmap[:, items.index("chunk_pts")] = chunk_pts
        full_cumsum = np.empty((n + 1,), dtype=np.int32)
        np.cumsum(chunk_pts, out=full_cumsum[1:])
        full_cumsum[0] = 0
        mmap[:, items.index("chunk1d_begin")] = full_cumsum[:-1]
        mmap[:, items.index("chunk1d_end")] = full_cumsum[1:]
        mmap.flush()
        del mmap

    def get_report_memmap(self, calc_name, mode='r+'):
        return open_memmap(filename=self.report_path(calc_name), mode=mode)

    def init_data_mmaps(self, calc_name):
        """ core.py:2362-2425 """
        state = self._calc_data[calc_name]["state"]
        data_type = {"Z": state.complex_type, "U": self.int_type,
                     "stop_reason": self.termination_type,
                     "stop_iter": self.int_type}
        data_path = self.data_path(calc_name)
        pts_count = self.pts_count(calc_name)
        f_complex, f_int, stop_codes = self.saved_codes(state.codes)
        data_dim = {"Z": (len(f_complex), pts_count), "U": (len(f_int), pts_count),
                    "stop_reason": (1, pts_count), "stop_iter": (1, pts_count)}
        for key in self.SAVE_ARRS:
            mmap = open_memmap(filename=data_path[key], mode='w+',
                               dtype=np.dtype(data_type[key]),
                               shape=data_dim[key], fortran_order=False)
            del mmap
        subset = state.subset
        if subset is not None:
            mmap = open_memmap(filename=data_path["subset"], mode='w+',
                               dtype=bool, shape=(self.ny * self.nx,),
                               fortran_order=False)
            beg = 0
            for chunk_slice in self.chunk_slices():
                (ix, ixx, iy, iyy) = chunk_slice
                end = beg + (ixx - ix) * (iyy - iy)
                mmap[beg:end] = subset[chunk_slice]
                beg = end
            del mmap

    def get_data_memmap(self, calc_name, key, mode='r+'):
        return open_memmap(filename=self.data_path(calc_name)[key], mode=mode)

    def _stored_rows(self, calc_name):
        state = self._calc_data[calc_name]["state"]
        (complex_codes, int_codes, stop_codes) = state.codes
        f_complex = self.filter_stored_codes(complex_codes)
        f_int = self.filter_stored_codes(int_codes)
        return {"Z": [complex_codes.index(c) for c in f_complex],
                "U": [int_codes.index(c) for c in f_int],
                "stop_reason": [0], "stop_iter": [0]}

    def update_data_mmaps(self, calc_name, chunk_slice, Z, U, stop_reason,
                          stop_iter):
        """ core.py:2437-2472 """
        report = self.get_report_memmap(calc_name, mode="r")
        rank = self.chunk_rank(chunk_slice)
        beg, end = int(report[rank, 0]), int(report[rank, 1])
        rows = self._stored_rows(calc_name)
        arrs = {"Z": Z, "U": U, "stop_reason": stop_reason, "stop_iter": stop_iter}
        for key in self.SAVE_ARRS:
            mmap = self.get_data_memmap(calc_name, key, mode="r+")
            for field, f_field in enumerate(rows[key]):
                mmap[field, beg:end] = arrs[key][f_field, :]
            mmap.flush()

    def update_report_mmap(self, calc_name, chunk_slice, stop_reason=None):
        mmap = self.get_report_memmap(calc_name, mode="r+")
        mmap[self.chunk_rank(chunk_slice), self.REPORT_ITEMS.index("done")] = 1
        mmap.flush()
        del mmap

    def reload_data(self, chunk_slice, calc_name):
        """ core.py:2475-2512 """
        report = self.get_report_memmap(calc_name, mode="r")
        rank = self.chunk_rank(chunk_slice)
        beg, end = int(report[rank, 0]), int(report[rank, 1])
        arr = {}
        for key in self.SAVE_ARRS:
            arr[key] = self.get_data_memmap(calc_name, key, mode="r")[:, beg:end]
        state = self._calc_data[calc_name]["state"]
        c_pix = np.ravel(self.chunk_pixel_pos(chunk_slice, False, None))
        chunk_subset = None
        if state.subset is not None:
            chunk_subset = np.asarray(state.subset[chunk_slice], dtype=bool)
            c_pix = c_pix[chunk_subset]
        return (chunk_subset, c_pix, arr["Z"], arr["U"], arr["stop_reason"],
                arr["stop_iter"])

    def clean_up(self, calc_name=None):
        """ core.py:1568-1614 """
        if calc_name is None:
            calc_name = "*"
        patterns = (calc_name + "_*.arr", calc_name + ".report",
                    calc_name + ".fingerprint", "ref_pt.dat")
        data_dir = os.path.join(self.directory, "data")
        if not os.path.isdir(data_dir):
            return
        for pattern in patterns:
            with os.scandir(data_dir) as it:
                for entry in it:
                    if fnmatch.fnmatch(entry.name, pattern):
                        os.unlink(entry.path)
        for temp_attr in ("_FP_params", "_Zn_path"):
            if hasattr(self, temp_attr):
                delattr(self, temp_attr)

    # -- the tile loop ---------------------------------------------------------
    def _staging(self, npts, n_Z, n_U, complex_type):
        """ Page-locked host staging buffers for one batch (kept and reused). """
        key = (n_Z, max(n_U, 1), np.dtype(complex_type).str)
        cur = getattr(self, "_staging_bufs", None)
        if cur is not None and cur["key"] == key and cur["cap"] >= npts:
            if cur["cap"] == npts:
                return cur
        if cur is not None and (cur["key"] != key or cur["cap"] < npts):
            self._free_staging()
            cur = None
        if cur is None:
            cur = {"key": key, "cap": npts,
                   "c_pix": _native.pinned_empty((npts,), np.complex128),
                   "Z": _native.pinned_empty((n_Z, npts), complex_type),
                   "U": _native.pinned_empty((max(n_U, 1), npts), np.int32),
                   "stop_reason": _native.pinned_empty((1, npts), np.int8),
                   "stop_iter": _native.pinned_empty((1, npts), np.int32)}
            self._staging_bufs = cur
        return cur

    def _free_staging(self):
        cur = getattr(self, "_staging_bufs", None)
        if cur is not None:
            for k in ("c_pix", "Z", "U", "stop_reason", "stop_iter"):
                _native.pinned_free(cur[k])
            self._staging_bufs = None

    def __del__(self):
        try:
            self._free_staging()
        except Exception:
            pass

    def prepare_mmaps(self, calc_name):
        """ Create the report / data memmaps of a calculation now (what calc_raw
        does on its first call).  Multi-rank runs: rank 0 only, before a barrier
        (multi.calc_raw_sharded). """
        self.init_report_mmap(calc_name)
        self.init_data_mmaps(calc_name)
        self._calc_data[calc_name]["need_new_mmap"] = False

    def bind_mmaps(self, calc_name):
        """ Use the memmaps another rank has created (after a barrier). """
        for key in ["report"] + self.SAVE_ARRS:
            path = (self.report_path(calc_name) if key == "report"
                    else self.data_path(calc_name)[key])
            open_memmap(filename=path, mode="r+")        # raises if missing
        self._calc_data[calc_name]["need_new_mmap"] = False

    def calc_raw(self, calc_name, tile_validator=None):
        """ core.py:2724-2736 """
        if self._calc_data[calc_name]["need_new_mmap"]:
            if (getattr(tile_validator, "world", 1) > 1
                    and not getattr(tile_validator, "files_ready", False)):
                # creating truncates: a second rank doing it would wipe the
                # `done` flags and slabs the first one has written
                raise RuntimeError(
                    "calc_raw with a multi-rank tile_validator: the memmaps must be created "
                    "by one rank before the others open them -- use "
                    "fractalshades_b200.multi.calc_raw_sharded (rank 0 creates, barrier, "
                    "the others bind)")
            self.prepare_mmaps(calc_name)
        self.compute_rawdata_dev(calc_name, chunk_slice=None,
                                 tile_validator=tile_validator)

    def compute_rawdata_dev(self, calc_name, chunk_slice=None,
                            tile_validator=None):
        """ GPU tile scheduler (replaces the thread-pool loop of
        core.py:2515-2554).  Pending tiles are grouped into batches of
        consecutive chunk ranks; each batch is one C-ABI call whose outputs
        land directly in the contiguous slab range of the memmaps. """
        if chunk_slice is not None:
            tiles = [chunk_slice]
        else:
            tiles = list(self.chunk_slices())
        report = self.get_report_memmap(calc_name, mode="r")
        done_col = self.REPORT_ITEMS.index("done")
        pending = []
        for cs in tiles:
            if tile_validator is not None and not tile_validator(cs):
                continue
            rank = self.chunk_rank(cs)
            if report[rank, done_col] > 0:      # resume: finished tiles skipped
                continue
            pending.append((rank, cs))
        pending.sort()
        indep = self._calc_data[calc_name]["cycle_indep_args"]
        rows = self._stored_rows(calc_name)
        state = self._calc_data[calc_name]["state"]
        n_Z, n_U = len(state.codes[0]), len(state.codes[1])
        stats_acc = {}
        batch, batch_pts = [], 0

        def flush(batch):
            if not batch or self.is_interrupted():
                return
            sizes = [int(report[rank, self.REPORT_ITEMS.index("chunk_pts")])
                     for rank, cs in batch]
            npts = int(sum(sizes))
            if npts == 0:
                return
            # page-locked staging buffers (reused across batches / frames): the
            # H2D / D2H copies of the slab pipeline then run at full PCIe speed
            bufs = self._staging(npts, n_Z, n_U, state.complex_type)
            Z = bufs["Z"][:, :npts] if npts == bufs["cap"] else None
            if state.subset is None:
                # full tiles: only their axes cross PCIe, the pixel grid is
                # built on the device (fsb_*_run_grid)
                c_pix = TileAxes(self, [cs for rank, cs in batch])
                shapes = None
            else:
                c_pix = bufs["c_pix"][:npts]
                off = 0
                shapes = None
                for (rank, cs), n in zip(batch, sizes):
                    pix = np.ravel(self.chunk_pixel_pos(cs, False, None))
                    pix = pix[np.asarray(state.subset[cs], dtype=bool)]
                    c_pix[off:off + n] = pix
                    off += n
            if Z is None:        # batch smaller than the staging capacity
                Z = np.zeros([n_Z, npts], dtype=state.complex_type)
                U = np.zeros([n_U, npts], dtype=self.int_type)
                stop_reason = - np.ones([1, npts], dtype=self.termination_type)
                stop_iter = np.zeros([1, npts], dtype=self.int_type)
            else:
                U = bufs["U"][:n_U]
                stop_reason, stop_iter = bufs["stop_reason"], bufs["stop_iter"]
            ret = self.numba_cycle_call((c_pix, Z, U, stop_reason, stop_iter), indep,
                                        tiles=shapes)
            for k, v in (Fractal._last_stats or {}).items():
                stats_acc[k] = stats_acc.get(k, 0) + v
            if ret == self.USER_INTERRUPTED:
                return
            arrs = {"Z": Z, "U": U, "stop_reason": stop_reason,
                    "stop_iter": stop_iter}
            mm = {k: self.get_data_memmap(calc_name, k, mode="r+")
                  for k in self.SAVE_ARRS}
            rep = self.get_report_memmap(calc_name, mode="r+")
            # tiles of consecutive ranks are consecutive slabs: group them into
            # contiguous runs and copy each run of each field in one go, the
            # copies spread over a few threads (numpy releases the GIL)
            runs, off = [], 0
            for (rank, cs), n in zip(batch, sizes):
                beg, end = int(rep[rank, 0]), int(rep[rank, 1])
                assert end - beg == n
                if runs and runs[-1][1] == beg:
                    runs[-1][1] = end
                else:
                    runs.append([beg, end, off])
                off += n
            # The slabs go to the files with pwrite: filling fresh page-cache pages
            # through a write call costs half of what page-faulting them through the
            # mapping does (265 MB, 8 threads: 73 ms against 143 ms); the memmaps of
            # the readers share the same page cache.
            jobs = []
            step = 1 << 20
            fds = {key: os.open(mm[key].filename, os.O_RDWR) for key in self.SAVE_ARRS}
            for beg, end, o in runs:
                for key in self.SAVE_ARRS:
                    item, ncol, base = mm[key].itemsize, mm[key].shape[1], mm[key].offset
                    for field, f_field in enumerate(rows[key]):
                        for a0 in range(0, end - beg, step):
                            a1 = min(a0 + step, end - beg)
                            jobs.append((fds[key], base + (field * ncol + beg + a0) * item,
                                         arrs[key], f_field, o + a0, o + a1))

            def copy(job):
                fd, pos, src, f_field, s0, s1 = job
                buf = np.ascontiguousarray(src[f_field, s0:s1]).view(np.uint8)
                done = 0
                while done < buf.size:
                    done += os.pwrite(fd, buf[done:], pos + done)
            try:
                if len(jobs) > 4 and settings.enable_multithreading:
                    with concurrent.futures.ThreadPoolExecutor(
                            max_workers=min(settings.io_threads, os.cpu_count() or 1)) as pool:
                        list(pool.map(copy, jobs))
                else:
                    for job in jobs:
                        copy(job)
            finally:
                for fd in fds.values():
                    os.close(fd)
            for rank, cs in batch:
                rep[rank, done_col] = 1
            rep.flush()

        for rank, cs in pending:
            n = int(report[rank, self.REPORT_ITEMS.index("chunk_pts")])
            if batch and batch_pts + n > settings.gpu_batch_pts:
                flush(batch)
                batch, batch_pts = [], 0
            batch.append((rank, cs))
            batch_pts += n
        flush(batch)
        self.last_stats = stats_acc

    def evaluate_rawdata_final(self, calc_name, chunk_slice, postproc_options):
        """ core.py:2570-2592 : final render, tile computed on the fly """
        jitter = float(postproc_options.get("jitter", 0.))
        supersampling = postproc_options.get("supersampling", None)
        if isinstance(supersampling, str):
            supersampling = {"None": None}.get(supersampling,
                                               int(supersampling.split("x")[0])
                                               if "x" in supersampling else None)
        (cycle_dep_args, chunk_subset) = self.get_cycling_dep_args(
            calc_name, chunk_slice, final=True, jitter=jitter,
            supersampling=supersampling)
        indep = self._calc_data[calc_name]["cycle_indep_args"]
        tiles = None
        if chunk_subset is None:
            (ix, ixx, iy, iyy) = chunk_slice
            ss = supersampling or 1
            tiles = [((ixx - ix) * ss, (iyy - iy) * ss)]
        ret_code = self.numba_cycle_call(cycle_dep_args, indep, tiles=tiles)
        if ret_code == self.USER_INTERRUPTED:
            return None
        (c_pix, Z, U, stop_reason, stop_iter) = cycle_dep_args
        return (chunk_subset, c_pix, Z, U, stop_reason, stop_iter)


class TileAxes:
    """ Pixel offsets of a list of full tiles as per-tile axes (`tile_axes`):
    passed in place of `c_pix` to `numba_cycle_call`, the library then builds
    the pixel grid on the device (fsb_*_run_grid) -- 0.7 MB instead of 133 MB
    over PCIe for a 4K frame.  Not in the reference: the GPU tile scheduler's
    own argument. """

    def __init__(self, fractal, chunk_slices):
        ax, shapes = [], []
        for cs in chunk_slices:
            x, y = fractal.tile_axes(cs)
            ax += [x, y]
            shapes.append((x.shape[0], y.shape[0]))
        self.shapes = shapes
        self.axes = np.ascontiguousarray(np.concatenate(ax), dtype=np.float64)
        self.npts = int(sum(w * h for w, h in shapes))
        self.tw, self.th = tile_shape_arrays(shapes, self.npts)


def check_outputs(npts, Z, U, stop_reason, stop_iter, nz, c_pix):
    """ The seam trusts no caller for buffer sizes: the library copies
    nz * npts elements into Z, npts into the others. """
    def bad(name, a, shape, dtype=None):
        return ValueError(f"{name}: expected C-contiguous {shape}"
                          + (f" {np.dtype(dtype).name}" if dtype else "")
                          + f", got {getattr(a, 'shape', None)} {getattr(a, 'dtype', None)}")
    if c_pix is not None and (c_pix.dtype != np.complex128 or c_pix.ndim != 1
                              or not c_pix.flags["C_CONTIGUOUS"]):
        raise bad("c_pix", c_pix, (npts,), np.complex128)
    if (Z.ndim != 2 or Z.shape[0] != nz or Z.shape[1] != npts
            or not Z.flags["C_CONTIGUOUS"] or Z.dtype not in (np.complex128, np.float64)):
        raise bad("Z", Z, (nz, npts))
    if U is not None and (U.dtype != np.int32 or U.ndim != 2 or U.shape[0] < 1
                          or U.shape[1] != npts or not U.flags["C_CONTIGUOUS"]):
        raise bad("U", U, (">=1", npts), np.int32)
    if (stop_reason.dtype != np.int8 or stop_reason.size != npts
            or not stop_reason.flags["C_CONTIGUOUS"]):
        raise bad("stop_reason", stop_reason, (1, npts), np.int8)
    if (stop_iter.dtype != np.int32 or stop_iter.size != npts
            or not stop_iter.flags["C_CONTIGUOUS"]):
        raise bad("stop_iter", stop_iter, (1, npts), np.int32)


def tile_shape_arrays(tiles, npts):
    """ [(width, height), ...] -> two int32 arrays for the *_run_tiles calls """
    t = np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(-1, 2))
    tw = np.ascontiguousarray(t[:, 0])
    th = np.ascontiguousarray(t[:, 1])
    if int(np.sum(tw.astype(np.int64) * th)) != int(npts):
        raise ValueError("tiles do not cover the point list: "
                         f"{int(np.sum(tw.astype(np.int64) * th))} != {npts}")
    return tw, th


def _picklable(fp):
    """ fingerprints may hold mpmath numbers / objects: store a STABLE text for them.
    A default `repr` carries the object's address and would never match on reload
    (every run would silently recompute): objects are described by their class and
    their primitive public attributes instead (a projection: its constructor
    parameters; a subset array: the calculation and field it was built from). """
    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, (int, float, str, bool, type(None))):
            return v
        if isinstance(v, (list, tuple)):
            return repr([conv(x) for x in v])
        r = repr(v)
        if " at 0x" not in r:
            return r
        attrs = {}
        for k, x in sorted(getattr(v, "__dict__", {}).items()):
            if k.startswith("_") or k == "fractal":
                continue
            if isinstance(x, (int, float, str, bool, type(None))):
                attrs[k] = x
            elif isinstance(x, (complex, np.generic)) or type(x).__module__.startswith("mpmath"):
                attrs[k] = repr(x)
        return f"{type(v).__module__}.{type(v).__qualname__}{attrs}"
    return conv(fp)
