# -*- coding: utf-8 -*-
"""
Exponential-map database of a deep zoom, rendered on a B200 -- the first half of
the reference's movie pipeline (examples/movies/with_DEM/zoom_script_DEM.py:
`plotter.save_db(relpath="expmap.postdb", ...)` on an `Expmap` projection).

The h axis of the map covers 55 decades; it is walked in steps
(`Db_writer.save_db` -> set_exp_zoom_step + reset_bla_tree per step, as
Fractal_plotter.save_expdb_by_steps does), every step one fused GPU call:
pixel kernels + continuous iteration / DEM / normal / field lines / Blinn
coefficients; the files have the reference's `.db` / `.postdb` layout.
Needs a CUDA device: there is no CPU fallback.

    python examples/expmap_db.py [out_dir]
"""
import os
import sys
import time

import numpy as np
from numpy.lib.format import open_memmap

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractalshades_b200 as fs                      # noqa: E402
import fractalshades_b200.models as fsm              # noqa: E402
from fractalshades_b200 import db as fdb             # noqa: E402
from fractalshades_b200 import postproc as fpp       # noqa: E402
from fractalshades_b200 import projection            # noqa: E402


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "expmap_out"
    fs.settings.no_newton = True
    f = fsm.Perturbation_mandelbrot(out_dir)
    f.zoom(precision=70,
           x="-1.929319698524937920226708049698305350754670432084006734339806946",
           y="-0.0000000000000000007592779387989739090287550144163328879329853232537252481600401185",
           dx="7.032184999234219e-55", nx=1600, xy_ratio=1.0, theta_deg=0.,
           projection=projection.Expmap(hmin=0., hmax=127.5, rotates_df=False,
                                        orientation="horizontal"))
    f.calc_std_div(calc_name="div", subset=None, max_iter=20000, M_divergence=1e3,
                   epsilon_stationnary=1e-3, BLA_eps=1e-6, interior_detect=False,
                   calc_dzndc=True)
    w = fdb.Db_writer(f, "div", fields=("cont_iter", "DEM", "normal"),
                      fieldlines=fpp.Fieldlines_pp(n_iter=4, swirl=0., endpoint_k=0.8))
    t0 = time.time()
    path = w.save_db(relpath="expmap.db", recovery_mode=True)
    t1 = time.time()
    db = open_memmap(path, mode="r")
    print(f"{path}: {db.shape} {db.dtype}, {w.n_steps} steps, {t1 - t0:.2f} s "
          f"({sum(s['kernel_ms'] for s in w.last_stats):.1f} ms in the pixel kernels)")
    for name, plane in zip(w.postnames, db):
        print(f"  {name:12s} finite {np.isfinite(plane).mean() * 100:5.1f} %  "
              f"median {np.nanmedian(plane):.4g}")
    # one layer frozen as pixels, as the movie scripts do
    layer = fdb.Grey_layer("cont_iter", func=np.log, probes_z=(2., 9.),
                           colors=[(0.05, 0.05, 0.2), (0.9, 0.7, 0.2), (1., 1., 1.)],
                           mask_color=(0.1, 0.1, 0.1))
    p = w.save_db(relpath="expmap.postdb", postdb_layer=layer, recovery_mode=False)
    px = open_memmap(p, mode="r")
    print(f"{p}: {px.shape} {px.dtype}")


if __name__ == "__main__":
    main()
