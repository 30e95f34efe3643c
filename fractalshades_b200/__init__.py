# -*- coding: utf-8 -*-
"""
fractalshades_b200 -- B200 (sm_100a) implementation of the per-pixel iteration
hot path of GBillotey/Fractalshades, behind the reference's own Python API:

    import fractalshades_b200 as fs
    import fractalshades_b200.models as fsm
    f = fsm.Perturbation_mandelbrot(directory)
    f.zoom(precision=..., x=..., y=..., dx=..., nx=..., xy_ratio=..., theta_deg=0.)
    f.calc_std_div(calc_name=..., subset=None, max_iter=..., ...)
    f.calc_raw(calc_name)          # GPU tile scheduler -> reference memmaps

The pixel loops run in hand-written CUDA (fractalshades_b200/csrc) reached
through a ctypes C ABI (include/fsb200.h).  No torch, no Triton, no numba, and
no CPU fallback: compute calls raise if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"

from . import settings
from . import projection
from . import xrange
from .core import Fractal, KernelSpec, USER_INTERRUPTED
from .perturbation import PerturbationFractal, FrameHandle, create_frame
from . import models
from .views import VIEWS
