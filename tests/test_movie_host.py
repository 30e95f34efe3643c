# -*- coding: utf-8 -*-
"""
CPU test of the zoom-movie host logic (BASELINE config 5): geometric widths,
per-frame precision, and ONE reference orbit (computed for the deepest frame)
reused by every shallower frame through the ref-point matching rule of the
reference (perturbation.py:211-253).  No GPU: the per-frame tables are built
with `frame_tables()` (host only).
"""
import os
import tempfile

import mpmath
import numpy as np

import fractalshades_b200 as fsb
import fractalshades_b200.models as fsm
from fractalshades_b200 import movie, multi


def _bind(f, calc_kwargs):
    kw = dict(calc_name="movie", subset=None, **calc_kwargs)
    ret = type(f).calc_std_div.__wrapped__(f, **kw)
    for k, v in kw.items():
        setattr(f, k, v)
    ret["set_state"]()(f)
    f._kernel_options = vars(ret["iterate"]()).copy()


def test_zoom_sequence_shares_one_orbit():
    v = fsb.VIEWS["deep_julia_2608"]
    d = tempfile.mkdtemp()
    calc = dict(max_iter=3000, M_divergence=1e3, epsilon_stationnary=1e-3, BLA_eps=1e-6,
                interior_detect=False, calc_dzndc=True)
    seq = movie.ZoomSequence(fsm.Perturbation_mandelbrot, d, x=v["x"][:400], y=v["y"][:400],
                             dx_start="1e-10", dx_end="1e-340", n_frames=6, nx=64,
                             xy_ratio=16 / 9., precision=370, calc_kwargs=calc)
    w = [float(mpmath.log10(x)) for x in seq.widths]
    assert np.allclose(np.diff(w), -66.0) and w[0] == -10.0
    assert movie.required_precision(seq.widths[0], 64) < 30 < movie.required_precision(seq.widths[-1], 64)
    seq.prepare_orbit(rank=0)
    ref_file = os.path.join(d, "data", "ref_pt.dat")
    mtime = os.path.getmtime(ref_file)
    seen_xr = set()
    for k in range(6):
        f = seq._fractal(k)
        _bind(f, calc)
        t = f.frame_tables()
        assert os.path.getmtime(ref_file) == mtime          # orbit reused, not recomputed
        assert len(t["Zn_path"]) == 3001
        assert t["xr_detect"] == (w[k] < -300)
        seen_xr.add(t["xr_detect"])
        # per-frame scalars follow dx
        assert abs(np.log2(t["lin_scale"]) + t["lin_scale_e"] - w[k] * np.log2(10)) < 1e-6
        assert t["kc"] > 0
    assert seen_xr == {False, True}
    # one directory per frame (ranks never write the same parameter /
    # fingerprint / report file), one shared orbit cache
    dirs = {seq._fractal(k).directory for k in range(6)}
    assert len(dirs) == 6 and all(os.path.dirname(p) == d for p in dirs)
    assert {seq._fractal(k).ref_point_file() for k in range(6)} == {ref_file}
    assert sorted(sum((multi.frames_for_rank(6, r, 4) for r in range(4)), [])) == list(range(6))
