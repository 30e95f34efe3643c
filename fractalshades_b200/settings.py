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
# False (the reference's default): the reference point of a perturbation frame
# is the nucleus found by the ball method + Newton descent around the image
# centre (every model: holomorphic power 2 and power N, burning-ship family).
# True: always the
# image centre -- what bench.py and the parity fixtures use.
no_newton = False              # settings.py:25
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
# Threads writing the raw planes of a batch of tiles to the memmap files (pwrite).
io_threads = 16
