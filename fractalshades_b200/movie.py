# -*- coding: utf-8 -*-
"""
Deep-zoom frame sequences (BASELINE config 5): N independent frames on one
centre, dx going geometrically from dx_start to dx_end.  This is the hot-path
view of the reference's `Custom_sequence` movies, which render every frame from
scratch through `plotter.save_db` (movie/movie.py:404-470): here each frame is
one pass of the per-pixel path.

* The reference orbit is computed ONCE, for the deepest frame (highest
  precision, largest max_iter), and reused by every shallower frame through the
  reference's own ref-point matching rule (perturbation.py:211-253).
* Per frame only the dx-dependent tables are rebuilt on the GPU (dZndc path
  scan, BLA tree: `kc` and the derivative scale change with dx).
* Frames are independent units: `multi.frames_for_rank` deals them to the
  ranks (one process per GPU), no collective on the data path.
"""
import os
import time

import mpmath
import numpy as np

from . import multi
from . import _native


def frame_widths(dx_start, dx_end, n_frames, precision):
    """ geometric sequence of image widths, as mpmath numbers """
    mpmath.mp.dps = precision
    a, b = mpmath.mpf(dx_start), mpmath.mpf(dx_end)
    if n_frames == 1:
        return [b]
    la, lb = mpmath.log(a), mpmath.log(b)
    return [mpmath.exp(la + (lb - la) * k / (n_frames - 1)) for k in range(n_frames)]


def required_precision(dx, nx, margin=10):
    """ digits needed for a frame of width dx (perturbation.py:123-135) """
    with mpmath.workdps(20):
        return int(-mpmath.log10(mpmath.mpf(dx) / nx / nx)) + margin


class ZoomSequence:
    def __init__(self, model_cls, directory, *, x, y, dx_start, dx_end, n_frames,
                 nx, xy_ratio, precision, calc_kwargs, model_kwargs=None,
                 theta_deg=0., zoom_kwargs=None):
        self.model_cls = model_cls
        self.directory = directory
        self.x, self.y = x, y
        self.n_frames = n_frames
        self.nx, self.xy_ratio, self.theta_deg = nx, xy_ratio, theta_deg
        self.precision = precision
        self.calc_kwargs = dict(calc_kwargs)
        self.model_kwargs = dict(model_kwargs or {})
        self.zoom_kwargs = dict(zoom_kwargs or {})
        self.widths = frame_widths(dx_start, dx_end, n_frames, precision)

    def frame_dir(self, k):
        return os.path.join(self.directory, f"frame_{k:04d}")

    def _fractal(self, k, directory=None):
        """ Frame k lives in its own directory (parameter / fingerprint / report
        files and memmaps are per frame: ranks never write the same file); the
        orbit cache is the shared one under `self.directory`. """
        f = self.model_cls(directory or self.frame_dir(k), **self.model_kwargs)
        f.ref_point_dir = self.directory
        dx = self.widths[k]
        prec = min(self.precision, max(required_precision(dx, self.nx), 20))
        # zoom() sets mpmath.mp.dps: parse the centre at full precision first
        mpmath.mp.dps = self.precision
        x, y = mpmath.mpf(self.x), mpmath.mpf(self.y)
        f.zoom(precision=prec, x=x, y=y, dx=dx, nx=self.nx, xy_ratio=self.xy_ratio,
               theta_deg=self.theta_deg, **self.zoom_kwargs)
        return f

    def prepare_orbit(self, rank=0, wait_s=3600.):
        """ deepest frame first: its orbit serves all the others.  With several
        ranks sharing `directory`, rank 0 computes, the others wait for it. """
        marker = os.path.join(self.directory, "data", "ref_pt.ready")
        if rank == 0:
            deepest = int(np.argmin([float(mpmath.log10(w)) for w in self.widths]))
            f = self._fractal(deepest)
            f.precision_used = mpmath.mp.dps
            for k, v in dict(calc_name="movie", subset=None, **self.calc_kwargs).items():
                setattr(f, k, v)
            t0 = time.time()
            mpmath.mp.dps = self.precision
            f.get_FP_orbit()
            os.makedirs(os.path.dirname(marker), exist_ok=True)
            open(marker, "w").write("ok")
            return time.time() - t0
        t0 = time.time()
        while not os.path.exists(marker):
            if time.time() - t0 > wait_s:
                raise RuntimeError("timed out waiting for the reference orbit")
            time.sleep(0.2)
        return 0.

    def render(self, rank=0, world=1, store=False, on_frame=None, pp_fields=None):
        """ Render this rank's frames.  store=True writes the reference-layout
        memmaps under <directory>/frame_XXXX/; store=False keeps the outputs in
        page-locked staging buffers only (benchmarks): the raw planes, or --
        pp_fields, e.g. ("cont_iter", "DEM") -- the post-processed float32
        fields of the fused call (postproc.frame_fields: 9 B per point leave
        the device instead of 41). """
        out = []
        for k in multi.frames_for_rank(self.n_frames, rank, world):
            t0 = time.time()
            f = self._fractal(k)
            f.calc_std_div(calc_name="movie", subset=None, **self.calc_kwargs)
            t1 = time.time()
            if store:
                f.calc_raw("movie")          # memmaps under frame_XXXX/data/
                st = f.last_stats
            elif pp_fields:
                from . import postproc
                _, st = postproc.frame_fields(f, "movie", fields=pp_fields, copy=False)
            else:
                st = render_frame_to_staging(f, "movie")
            t2 = time.time()
            rec = {"frame": k, "dx": mpmath.nstr(self.widths[k], 6), "setup_s": t1 - t0,
                   "render_s": t2 - t1, "kernel_ms": st.get("kernel_ms", 0.),
                   "sum_stop_iter": int(st.get("sum_stop_iter", 0)),
                   "xr": bool(getattr(f, "xr_detect_activated", False))}
            f._release_indep_args(f._calc_data["movie"]["cycle_indep_args"])
            f._free_staging()
            out.append(rec)
            if on_frame:
                on_frame(rec)
        return out


_STAGE = {}


def render_frame_to_staging(f, calc_name):
    """ One whole frame through the seam (numba_cycle_call) with page-locked
    host buffers that are reused from frame to frame; returns the stats. """
    state = f._calc_data[calc_name]["state"]
    indep = f._calc_data[calc_name]["cycle_indep_args"]
    n_Z, n_U = len(state.codes[0]), len(state.codes[1])
    npts = f.nx * f.ny
    key = (npts, n_Z, n_U, np.dtype(state.complex_type).str)
    if _STAGE.get("key") != key:
        for a in _STAGE.get("bufs", {}).values():
            _native.pinned_free(a)
        _STAGE["key"] = key
        _STAGE["bufs"] = {
            "Z": _native.pinned_empty((n_Z, npts), state.complex_type),
            "U": _native.pinned_empty((max(n_U, 1), npts), np.int32),
            "sr": _native.pinned_empty((1, npts), np.int8),
            "si": _native.pinned_empty((1, npts), np.int32)}
        _STAGE["grid"] = None
    b = _STAGE["bufs"]
    gkey = (f.nx, f.ny, f.xy_ratio)
    if _STAGE.get("grid") != gkey:       # the per-tile axes depend on the grid only
        from .core import TileAxes
        _STAGE["axes"] = TileAxes(f, list(f.chunk_slices()))
        _STAGE["grid"] = gkey
    rc = f.numba_cycle_call((_STAGE["axes"], b["Z"], b["U"][:n_U], b["sr"], b["si"]), indep)
    if rc != 0:
        raise RuntimeError("frame interrupted")
    from .core import Fractal
    return dict(Fractal._last_stats)
