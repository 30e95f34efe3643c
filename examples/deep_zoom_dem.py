# -*- coding: utf-8 -*-
"""
Deep-zoom frame with the reference's Python API, rendered on a B200.

Same calls as the reference's examples/batch_mode/11-run_perturbdeep.py
(zoom -> calc_std_div -> raw fields), plus the GPU post-processing of this
package.  Needs a CUDA device: there is no CPU fallback.

    python examples/deep_zoom_dem.py [out_dir]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractalshades_b200 as fs                      # noqa: E402
import fractalshades_b200.models as fsm              # noqa: E402
from fractalshades_b200 import postproc as fpp       # noqa: E402
from fractalshades_b200.views import VIEWS           # noqa: E402


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "deep_zoom_out"
    view = VIEWS["deep_julia_2608"]      # centre of the reference's example 11 (3 500 digits)
    # the image centre is the reference point (fs.settings.no_newton = False, the
    # reference's default, would first look for the nucleus of the minibrot)
    fs.settings.no_newton = True
    f = fsm.Perturbation_mandelbrot(out_dir)
    f.zoom(precision=270, x=view["x"][:290], y=view["y"][:290], dx="1e-250", nx=1920,
           xy_ratio=16 / 9., theta_deg=0.)
    t0 = time.time()
    f.calc_std_div(calc_name="div", subset=None, max_iter=1000000, M_divergence=1e3,
                   epsilon_stationnary=1e-3, BLA_eps=1e-6, interior_detect=False,
                   calc_dzndc=True)
    t1 = time.time()
    # (a) the reference's data flow: raw fields into the tile memmaps
    f.calc_raw("div")
    t2 = time.time()
    stop_iter = f.get_data_memmap("div", "stop_iter", mode="r")
    # (b) fused: continuous iteration / DEM / normal straight from the GPU
    fields, stats = fpp.frame_fields(f, "div")
    t3 = time.time()
    nu = fpp.to_image(f, fields["cont_iter"])
    esc = fpp.to_image(f, fields["stop_reason"]) == 1
    print(f"frame setup (orbit, dZndc scan, BLA tree): {t1 - t0:.2f} s")
    print(f"calc_raw -> memmaps: {t2 - t1:.3f} s   fused post-processing: {t3 - t2:.3f} s "
          f"(kernel {stats['kernel_ms']:.1f} ms)")
    print(f"{f.nx} x {f.ny} px, {int(np.sum(stop_iter, dtype=np.int64)):,} effective iterations, "
          f"{100 * esc.mean():.1f} % escaped, continuous iteration in "
          f"[{np.nanmin(nu[esc]):.1f}, {np.nanmax(nu[esc]):.1f}]")


if __name__ == "__main__":
    main()
