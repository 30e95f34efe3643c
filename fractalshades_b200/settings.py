# -*- coding: utf-8 -*-
"""
Module-level flags read by the hot path; same names and defaults as the
reference's `fractalshades.settings` (settings.py:7-82) where they exist.
"""
enable_multithreading = True   # settings.py:7 (host-side tile post-processing)
skip_calc = False              # settings.py:10
newton_zoom_level = 1.e-5      # settings.py:14 ; also gates BLA
std_zoom_level = 1.e-8         # settings.py:18
xrange_zoom_level = 1.e-300    # settings.py:22
# The ball-method / Newton nucleus search is not part of this hot path (SURVEY
# section 8 f-1): the reference point is always the image centre.
no_newton = True               # settings.py:25 (reference default: False)
inspect_calc = False           # settings.py:30
chunk_size = 200               # settings.py:34
BLA_compression = 3            # settings.py:40 (fixed: the kernels fold 3 levels)
postproc_dtype = "float32"     # settings.py:82

# ---- additions of this implementation ------------------------------------
# Use the -fmad=false build (bit-reproducible against the IEEE-strict oracle).
strict_ieee = False
# Maximum number of points submitted to the GPU in one launch by the tile
# scheduler (bounds device memory: ~60 B/pt for zn+dzndc).
gpu_batch_pts = 1 << 25
