# -*- coding: utf-8 -*-
"""
Database writer: the post-processed fields of a
frame saved in the reference's on-disk layout, and the stepped driver of the
large exponential maps its movie pipeline is built on.

Mirrors `Fractal_plotter.save_db` (core.py:812-889) and what it calls:

  <relpath>            `.db`     float32 memmap (n_posts, ny, nx), image row order
                                 (row 0 = top), one plane per post-processed field,
                                 "normal" as normal_x / normal_y (open_db :971-1050,
                                 push_db :1096-1121)
                       `.postdb` uint8 memmap (ny, nx, channels): ONE layer frozen
                                 as pixels (open_postdb :1126-1199, push_postdb
                                 :1221-1246)
  <root>_status<ext>   int32 (n_tiles,): tiles already written (open_db_status
                                 :1058-1093); with recovery_mode=True a run resumes
                                 from them
  <relpath>.info       the text description written at :866-887

  Expmap projection    `save_expdb_by_steps` (:891-952): the h axis is walked in steps
                       of `nt` pixels; each step sets the projection's step window
                       (`set_exp_zoom_step`), rebuilds the frame tables for it
                       (`reset_bla_tree`: BLA radii and derivative scale of the step)
                       and renders the tiles that END inside the step.

The fields come from the fused GPU call (`postproc.frame_fields`: pixel kernels +
post-processing, nothing but the fields crosses PCIe); the reference's loop over
`process()` -> `calc_raw` -> numpy post-processing per tile is what it replaces.
One difference, on purpose: the reference's `process()` re-creates (zeroes) the
database on every step unless recovery_mode=True -- which is why its movie scripts
always pass it; here the steps of ONE save_db call always accumulate.
"""
import datetime
import os

import numpy as np
from numpy.lib.format import open_memmap

from . import postproc as fpp
from . import projection as _projection


class Grey_layer:
    """ Minimal stand-in for the reference's colour layers in `.postdb` mode: one
    post-processed field mapped to uint8 pixels the way `Color_layer.child_crop`
    normalises it (colors/layers.py:571-583: optional function, rescale between
    the two probes, triangle-wave wrap), then a piecewise-linear colour ramp
    (`colors`: (n, 3) RGB stops in [0, 1]; None = one grey channel). """
    def __init__(self, postname, func=None, probes_z=(0., 1.), colors=None, mask_color=None):
        self.postname = postname
        self.func = func
        self.probes_z = tuple(float(v) for v in probes_z)
        self.colors = None if colors is None else np.asarray(colors, np.float64)
        self.mask_color = mask_color
        self.n_channels = 1 if colors is None else 3

    def pixels(self, arr, mask=None):
        arr = np.asarray(arr, np.float64)
        if self.func is not None:
            arr = self.func(arr)
        z0, z1 = self.probes_z
        with np.errstate(all="ignore"):
            arr = (arr - z0) / (z1 - z0)
            e = np.floor((arr + 1.) / 2.)
            arr = np.abs((arr - 2. * e) * (-1.) ** e)
        arr = np.nan_to_num(arr, nan=0., posinf=1., neginf=0.)
        if self.colors is None:
            px = np.uint8(arr * 255)[..., np.newaxis]
        else:
            stops = np.linspace(0., 1., len(self.colors))
            px = np.stack([np.interp(arr, stops, self.colors[:, c]) for c in range(3)], axis=-1)
            px = np.uint8(px * 255)
        if mask is not None and self.mask_color is not None:
            mc = np.uint8(np.asarray(self.mask_color, np.float64)[:self.n_channels] * 255)
            px[mask] = mc
        return px


class Db_writer:
    def __init__(self, fractal, calc_name, fields=("cont_iter", "DEM", "normal"), floor_iter=0,
                 px_snap=None, dtype=np.float32, fieldlines=None):
        self.fractal, self.calc_name = fractal, calc_name
        self.floor_iter, self.px_snap, self.fieldlines = floor_iter, px_snap, fieldlines
        self.post_dtype = np.dtype(dtype)
        self.fields = tuple(fields)
        names = []
        for k in self.fields:
            names += ["normal_x", "normal_y"] if k == "normal" else [k]
        if fieldlines is not None:
            names.append("fieldlines")
        self.postnames = names
        self.last_stats = []

    # -- paths, core.py:760-808 and :1052-1056 ---------------------------------
    def db_path(self, relpath=None):
        return os.path.normpath(os.path.join(self.fractal.directory, relpath or "layers.db"))

    @staticmethod
    def status_path(db_path):
        root, ext = os.path.splitext(db_path)
        return root + "_status" + ext

    @property
    def db_shape(self):
        return (self.fractal.ny, self.fractal.nx)

    # -- memmaps ---------------------------------------------------------------
    def _open_status(self, db_path, recover):
        n_chunk = self.fractal.chunks_count
        path = self.status_path(db_path)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        if recover:
            try:
                st = open_memmap(filename=path, mode="r+")
                if st.shape == (n_chunk,):
                    return st
            except (FileNotFoundError, ValueError):
                pass
        st = open_memmap(filename=path, mode="w+", dtype=np.int32, shape=(n_chunk,),
                         fortran_order=False, version=None)
        st[:] = 0
        return st

    def _open_db(self, db_path, shape, dtype, recover):
        """ returns (memmap, status); an existing database of the expected shape
        and type is kept in recovery mode, anything else is re-created and its
        status flags cleared (open_db / open_postdb) """
        status = self._open_status(db_path, recover)
        if recover:
            try:
                mm = open_memmap(filename=db_path, mode="r+")
                if mm.shape == shape and mm.dtype == np.dtype(dtype):
                    return mm, status
                del mm
            except (FileNotFoundError, ValueError):
                pass
        os.makedirs(os.path.dirname(db_path), exist_ok=True)
        mm = open_memmap(filename=db_path, mode="w+", dtype=np.dtype(dtype), shape=shape,
                         fortran_order=False, version=None)
        status[:] = 0
        return mm, status

    # -- tiles -----------------------------------------------------------------
    def _render(self, tiles):
        """ fields of the listed tiles: dict name -> 1-D tile-ordered array """
        f = self.fractal
        indep = f._calc_data[self.calc_name]["cycle_indep_args"]
        if indep[0] == "perturb":
            out, stats = fpp.frame_fields(f, self.calc_name, fields=self.fields,
                                          floor_iter=self.floor_iter, px_snap=self.px_snap,
                                          dtype=self.post_dtype, tiles=tiles, copy=False,
                                          fieldlines=self.fieldlines)
            self.last_stats.append(stats)
            return out
        # standard (double-precision) models: the raw seam with per-tile axes, then the
        # stand-alone post-processing call on the raw rows
        from .core import TileAxes
        state = f._calc_data[self.calc_name]["state"]
        axes = TileAxes(f, tiles)
        n_Z = len(state.codes[0])
        Z = np.zeros((n_Z, axes.npts), state.complex_type)
        U = np.zeros((len(state.codes[1]), axes.npts), np.int32)
        sr = -np.ones((1, axes.npts), np.int8)
        si = np.zeros((1, axes.npts), np.int32)
        rc = f.numba_cycle_call((axes, Z, U, sr, si), indep)
        if rc != 0:
            raise RuntimeError("frame interrupted")
        self.last_stats.append(dict(type(f)._last_stats or {}))
        c_pix = np.concatenate([np.ravel(f.chunk_pixel_pos(cs, False, None)) for cs in tiles])
        out = fpp.fields_from_raw(f, self.calc_name, Z, si, fields=self.fields,
                                  floor_iter=self.floor_iter, px_snap=self.px_snap,
                                  dtype=self.post_dtype, c_pix=c_pix, fieldlines=self.fieldlines)
        out["stop_reason"] = sr[0]
        return out

    def _push(self, mm, status, tiles, out, layer):
        f = self.fractal
        off = 0
        for cs in tiles:
            (ix, ixx, iy, iyy) = cs
            w, h = ixx - ix, iyy - iy
            n = w * h
            if layer is None:                       # push_db
                for p, name in enumerate(self.postnames):
                    mm[p, iy:iyy, ix:ixx] = out[name][off:off + n].reshape(h, w)
            else:                                   # push_postdb
                mask = (out["stop_reason"][off:off + n] != 1).reshape(h, w)
                mm[iy:iyy, ix:ixx, :] = layer.pixels(out[layer.postname][off:off + n].reshape(h, w), mask)
            status[f.chunk_rank(cs)] = 1
            off += n

    def _run(self, mm, status, layer, validator=None):
        f = self.fractal
        tiles = [cs for cs in f.chunk_slices()
                 if (validator is None or validator(cs)) and status[f.chunk_rank(cs)] == 0]
        if not tiles:
            return 0
        self._push(mm, status, tiles, self._render(tiles), layer)
        return len(tiles)

    # -- the stepped exponential map, core.py:891-952 ----------------------------
    def exp_steps(self):
        """ (r, stp, step_hmax, step_hmin) for each step of save_expdb_by_steps """
        from . import settings
        f = self.fractal
        proj = f.projection
        stp = proj.nt(f)
        hmin, hmax, nh = proj.hmin, proj.hmax, proj.nh(f)
        chunk_size = settings.chunk_size
        for r in range(0, nh + 1, stp):
            i_max = min(r + stp, nh)
            i_min = max(r - chunk_size, 0)
            yield (r, stp, (hmax * i_max + hmin * (nh - i_max)) / nh,
                   (hmax * i_min + hmin * (nh - i_min)) / nh)

    def _save_expdb_by_steps(self, mm, status, layer):
        f = self.fractal
        proj = f.projection
        horizontal = proj.orientation == "horizontal"
        n_steps = 0
        try:
            for (r, stp, step_hmax, step_hmin) in self.exp_steps():
                def validates(cs, r=r, stp=stp):
                    (_, ixx, _, iyy) = cs
                    return r < (ixx if horizontal else iyy) <= (r + stp)
                # the reference rebuilds the tables at every step; a step without a
                # pending tile changes nothing that is read later, so it is skipped
                if not any(validates(cs) and status[f.chunk_rank(cs)] == 0 for cs in f.chunk_slices()):
                    continue
                proj.set_exp_zoom_step(step_hmax, step_hmin)
                data = f._calc_data[self.calc_name]
                data["cycle_indep_args"] = f.reset_bla_tree(data["cycle_indep_args"])
                self._run(mm, status, layer, validates)
                n_steps += 1
        finally:
            proj.del_exp_zoom_step()
        return n_steps

    # -- public ----------------------------------------------------------------
    def save_db(self, relpath=None, postdb_layer=None, recovery_mode=False):
        """ core.py:812-889.  postdb_layer: a layer object (`Grey_layer`) whose
        `postname` is one of the fields -> `.postdb`; None -> `.db` with every field. """
        f = self.fractal
        if postdb_layer is not None:
            if postdb_layer.postname not in self.postnames:
                raise ValueError(f"layer field {postdb_layer.postname!r} is not computed")
            path = self.db_path(relpath or f"{postdb_layer.postname}.postdb")
            shape, dtype = self.db_shape + (postdb_layer.n_channels,), np.uint8
        else:
            path = self.db_path(relpath)
            shape, dtype = (len(self.postnames),) + self.db_shape, self.post_dtype
        mm, status = self._open_db(path, shape, dtype, recovery_mode)
        self.last_stats = []
        if isinstance(f.projection, _projection.Expmap) and hasattr(f, "reset_bla_tree"):
            # (the reference steps every Expmap; its stepping only re-derives the BLA tree
            # and the derivative scale of a perturbation frame -- a standard model has
            # neither, and one pass gives the same arrays)
            self.n_steps = self._save_expdb_by_steps(mm, status, postdb_layer)
        else:
            self._run(mm, status, postdb_layer)
            self.n_steps = 1
        mm.flush(); status.flush()
        del mm, status
        with open(path + ".info", "w+") as info:
            info.write("Db file description\n")
            info.write(f"written time: {datetime.datetime.now()}\n")
            info.write(f"recovery_mode: {recovery_mode}\n\n")
            info.write("*array description*\n")
            info.write(f"  dtype: {self.post_dtype}\n")
            info.write(f"  shape: {(len(self.postnames),) + self.db_shape}\n")
            info.write("  supersampling: None\n")
            info.write(f"  postdb_layer: {None if postdb_layer is None else postdb_layer.postname}\n\n")
            info.write("*fields description*\n")
            for pn in self.postnames:
                info.write(f"  {pn}\n")
        return path
