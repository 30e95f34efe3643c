# -*- coding: utf-8 -*-
"""
Projections (mirror of the reference's projection.py for the hot path).

The reference hands numba closures (`projection.f`, `dzndc_modifier`) to its
kernels inside `cycle_indep_args`; closures cannot cross a C ABI, so each
projection here reduces itself to the plain parameters of `fsb_proj_desc`
(include/fsb200.h) with `c_abi_desc()`:

  Cartesian                    identity                    projection.py:155-223
  Cartesian(expmap_seam=s)     + dz/dc modifier |pix + 1e-6| s        :205-219
  Expmap(hmin, hmax, ...)      pix -> exp(hmoy + pix_to_ht pix)       :226-500
                               + dz/dc modifier exp(Re(pix_to_ht pix) + hshift)

`Generic_mapping` (an arbitrary user function, projection.py:506-717) has no
parametric form and is refused -- there is no fallback path.

`Expmap.df` / `dfBS` (the rotation of the derivatives applied by the
reference's host post-processing, projection.py:375-453) belong to the
consumers of the raw fields and are exposed as numpy functions.
"""
import math

import mpmath
import numpy as np

from . import settings
from ._native import (FsbProjDesc, FSB_PROJ_CARTESIAN, FSB_PROJ_EXPMAP,
                      FSB_DZNDC_MOD_NONE, FSB_DZNDC_MOD_EXPMAP, FSB_DZNDC_MOD_SEAM)


class Projection:
    scale = 1.0

    @property
    def init_kwargs(self):
        """ projection.py:52-62 """
        import inspect
        return {p: getattr(self, p)
                for p in inspect.signature(self.__init__).parameters}

    def __eq__(self, other):
        return (other.__class__ == self.__class__
                and other.init_kwargs == self.init_kwargs)

    def fingerprint(self):
        """ hashable description stored in zoom_kwargs (same role as the
        projection object itself in the reference's fingerprint) """
        return (type(self).__name__,) + tuple(
            (k, str(v)) for k, v in sorted(self.init_kwargs.items()))

    def adjust_to_zoom(self, fractal):
        self.fractal = fractal

    def c_abi_desc(self):
        raise NotImplementedError(
            f"projection {type(self).__name__} has no parametric form and "
            "cannot cross the C ABI (no fallback path)")


class Cartesian(Projection):
    def __init__(self, expmap_seam=None):
        """ projection.py:157-182 """
        self.scale = 1.0
        self.expmap_seam = expmap_seam

    def bounding_box(self, xy_ratio):
        return 1., 1. / xy_ratio

    @property
    def min_local_scale(self):
        return 1.

    @property
    def has_dzndc_modifier(self):
        return self.expmap_seam is not None

    def c_abi_desc(self):
        d = FsbProjDesc()
        d.kind = FSB_PROJ_CARTESIAN
        if self.expmap_seam is not None:       # projection.py:205-219
            d.dzndc_modifier = FSB_DZNDC_MOD_SEAM
            d.mod_param = float(self.expmap_seam)
        return d


class Expmap(Projection):
    def __init__(self, hmin, hmax, rotates_df=True, orientation="horizontal"):
        """ projection.py:227-297 """
        if not (0 <= hmin < hmax):
            raise ValueError(
                "Provide hmin, hmax with:  0 <= hmin < hmax for Expmap")
        self.use_step = False
        self.rotates_df = rotates_df
        self.orientation = orientation
        self.premul_1j = {"horizontal": False, "vertical": True}[orientation]
        # The reference keeps hmin / hmax as Xrange scalars beyond exp(hmax) >
        # 1e300 (only used by its step bookkeeping); plain values serve here.
        self.hmin = hmin
        self.hmax = hmax
        self.hmoy = (hmin + hmax) * 0.5
        self.dh = hmax - hmin

    def nh(self, fractal):
        return fractal.ny if self.premul_1j else fractal.nx

    def nt(self, fractal):
        return fractal.nx if self.premul_1j else fractal.ny

    @property
    def pix_to_ht(self):
        """ projection.py:306-313 (same Python expression, same rounding) """
        dh = self.dh
        return (1j * dh * self.xy_ratio) if self.premul_1j else dh

    # -- steps of a large exponential zoom, projection.py:316-337 ------------
    def set_exp_zoom_step(self, exp_step_hmax, exp_step_hmin):
        self.exp_step_hmax = exp_step_hmax
        self.exp_step_hmin = exp_step_hmin
        self.exp_step_hmoy = 0.5 * (exp_step_hmax + exp_step_hmin)
        self.exp_step_dh = exp_step_hmax - exp_step_hmin
        self.use_step = True

    def del_exp_zoom_step(self):
        self.use_step = False

    def adjust_to_zoom(self, fractal):
        """ projection.py:339-360 : xy_ratio is imposed by dh = 2 pi xy_ratio """
        if self.premul_1j:
            xy_ratio = (np.pi * 2.) / self.dh
            nx = int(fractal.nx * xy_ratio + 0.5)
        else:
            xy_ratio = self.dh / (np.pi * 2.)
            nx = fractal.nx
        fractal.xy_ratio = self.xy_ratio = xy_ratio
        fractal.zoom_kwargs["xy_ratio"] = xy_ratio
        fractal.nx = nx
        fractal.zoom_kwargs["nx"] = nx
        self.fractal = fractal

    @property
    def hshift(self):
        """ projection.py:461-465 """
        return (self.hmoy - self.exp_step_hmoy) if self.use_step else self.hmoy

    @property
    def has_dzndc_modifier(self):
        return True

    @property
    def scale(self):
        """ projection.py:474-481 """
        return mpmath.exp(self.exp_step_hmoy if self.use_step else self.hmoy)

    def bounding_box(self, xy_ratio):
        """ projection.py:483-493 """
        w = mpmath.exp(self.exp_step_hmax if self.use_step else self.hmax)
        return w, w

    @property
    def min_local_scale(self):
        return mpmath.exp(self.hmin)

    def c_abi_desc(self):
        d = FsbProjDesc()
        d.kind = FSB_PROJ_EXPMAP
        d.hmoy = float(self.hmoy)
        k = complex(self.pix_to_ht)
        d.pix_to_ht[0], d.pix_to_ht[1] = k.real, k.imag
        d.dzndc_modifier = FSB_DZNDC_MOD_EXPMAP
        d.mod_param = float(self.hshift)
        return d

    # -- consumers' side (host post-processing), projection.py:375-453 -------
    def df(self, pix):
        """ derivative rotation / scaling applied by the reference's
        post-processing to holomorphic dz/dc fields """
        ht = self.pix_to_ht * np.asarray(pix)
        if self.use_step:
            return np.exp(1j * np.imag(ht)) if self.rotates_df else np.ones_like(ht.real)
        return np.exp(ht) if self.rotates_df else np.exp(np.real(ht))

    def dfBS(self, pix):
        ht = self.pix_to_ht * np.asarray(pix)
        h, t = np.real(ht), np.imag(ht)
        if self.use_step:
            if self.rotates_df:
                c, s = np.cos(t), np.sin(t)
                return c, -s, s, c
            one = np.ones_like(h)
            return one, 0. * one, 0. * one, one
        r = np.exp(h)
        if self.rotates_df:
            cr, sr = np.cos(t) * r, np.sin(t) * r
            return -sr, -cr, cr, -sr
        return r, 0. * r, 0. * r, r


class Generic_mapping(Projection):
    def __init__(self, f, df):
        raise NotImplementedError(
            "Generic_mapping takes arbitrary Python functions, which cannot "
            "cross the C ABI of the GPU path (no fallback)")
