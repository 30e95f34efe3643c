# -*- coding: utf-8 -*-
"""
gen_golden_expdb.py -- `.db` fixture of a stepped exponential-map database from
the LIVE reference (build container only; TEST INFRASTRUCTURE).

The reference's own flow: Fractal_plotter.save_db (core.py:812-889) ->
save_expdb_by_steps (:891-952: per step set_exp_zoom_step + reset_bla_tree, the
tiles of the step's h range through process(tile_validator)) -> push_db
(:1096-1121) into the (n_posts, ny, nx) float32 memmap `layers.db` with its
`layers_status.db` flags.  The fixture keeps that array, the status flags and
the tile size used; tests/test_gpu_db.py builds the same database with
fractalshades_b200.db.save_db on the GPU.

    python tools/gen_golden_expdb.py     -> tests/golden/expdb_<case>.npz
"""
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))

DB_CASES = {"p_M2_expmap_E55_horiz": 50, "p_BS_f1_expmap_E30": 40}     # case -> chunk_size


def main(argv):
    import ref_harness as rh
    from cases import CASES
    fs = rh.load_reference()
    from fractalshades.postproc import (Postproc_batch, Continuous_iter_pp, DEM_pp,
                                        DEM_normal_pp)
    from fractalshades.colors.layers import Virtual_layer
    from numpy.lib.format import open_memmap
    for name in (argv or DB_CASES):
        case = dict(CASES[name])
        chunk = DB_CASES[name]
        old_chunk = fs.settings.chunk_size
        fs.settings.chunk_size = chunk
        fs.settings.postproc_dtype = "float32"
        fs.settings.no_newton = True
        try:
            with tempfile.TemporaryDirectory() as wd:
                f = rh.make_fractal(case, wd)
                pb = Postproc_batch(f, "c")
                pb.add_postproc("cont_iter", Continuous_iter_pp())
                pb.add_postproc("DEM", DEM_pp())
                pb.add_postproc("normal", DEM_normal_pp(kind="potential"))
                plotter = fs.Fractal_plotter(pb)
                for pn in ("cont_iter", "DEM", "normal"):
                    plotter.add_layer(Virtual_layer(pn, func=None, output=False))
                # recovery_mode=True as the reference's movie scripts do: without it every step's
                # process() re-creates (zeroes) the database (open_db, core.py:985-1050)
                path = plotter.save_db(recovery_mode=True)
                db = np.array(open_memmap(path, mode="r"))
                root, ext = os.path.splitext(path)
                status = np.array(open_memmap(root + "_status" + ext, mode="r"))
                info = open(path + ".info").read()
                posts = list(plotter.postnames)
                proj = f.projection
                meta = {"case": name, "chunk_size": chunk, "posts": posts, "nx": f.nx, "ny": f.ny,
                        "relpath": os.path.relpath(path, f.directory), "n_steps": None,
                        "exp_zoom_step": int(proj.nt(f)), "nh": int(proj.nh(f)),
                        "reference": "GBillotey/Fractalshades v1.2.1 Fractal_plotter.save_db, "
                                     "fastmath as shipped, postproc_dtype float32"}
                out = os.path.join(REPO, "tests", "golden", f"expdb_{name}.npz")
                np.savez_compressed(out, db=db, status=status, info=info, meta=json.dumps(meta))
                print(name, posts, db.shape, db.dtype, "status", status.tolist(), "size",
                      os.path.getsize(out))
        finally:
            fs.settings.chunk_size = old_chunk


if __name__ == "__main__":
    main(sys.argv[1:])
