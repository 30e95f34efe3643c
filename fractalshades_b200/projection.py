# -*- coding: utf-8 -*-
"""
Projections.  Only `Cartesian` (identity, reference projection.py:155-223) can
cross the C ABI; arbitrary numba closures (Expmap, Generic_mapping) cannot and
are refused -- there is no fallback path.
"""


class Projection:
    scale = 1.0

    def adjust_to_zoom(self, fractal):
        self.fractal = fractal


class Cartesian(Projection):
    def __init__(self, expmap_seam=None):
        if expmap_seam is not None:
            raise NotImplementedError(
                "Cartesian(expmap_seam=...) needs the dzndc_modifier closure; "
                "not supported by the GPU path")
        self.scale = 1.0
        self.expmap_seam = None

    def bounding_box(self, xy_ratio):
        return 1., 1. / xy_ratio

    @property
    def min_local_scale(self):
        return 1.
