# -*- coding: utf-8 -*-
"""
gen_golden_pp.py -- post-processing fixtures from the LIVE reference (build
container only; TEST INFRASTRUCTURE).  For each listed case the reference runs
its own calculation and then its own post-processing objects
(postproc.Continuous_iter_pp :352-406, DEM_pp :684-731, DEM_normal_pp :572-628
through Fractal.postproc, core.py:2739-2778) with postproc_dtype = float64.
The fixture holds the raw arrays the post-processing consumed and the fields it
produced; tests/test_gpu_postproc.py feeds the same raw arrays to k_postproc.

    python tools/gen_golden_pp.py            -> tests/golden/pp_<case>.npz
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))

PP_CASES = ["std_M2_cfg1", "p_M2_E20", "p_M2_deep250", "p_BS_f1_E30_skew", "std_BS_f1"]


def main(argv):
    import ref_harness as rh
    from cases import CASES
    fs = rh.load_reference()
    fs.settings.postproc_dtype = "float64"
    from fractalshades.postproc import (Postproc_batch, Continuous_iter_pp, DEM_pp,
                                        DEM_normal_pp)
    names = argv or PP_CASES
    for name in names:
        case = CASES[name]
        out = rh.run_case(case, keep_tables=False)
        f = out["fractal"]
        codes = f._calc_data["c"]["saved_codes"][0]
        have_deriv = ("dzndc" in codes) or ("dxnda" in codes)
        pb = Postproc_batch(f, "c")
        pb.add_postproc("cont_iter", Continuous_iter_pp())
        if have_deriv:
            pb.add_postproc("DEM", DEM_pp())
            pb.add_postproc("normal", DEM_normal_pp(kind="potential"))
        posts = list(pb.posts.keys())
        chunks = []
        for cs in f.chunk_slices():
            post_array, subset = f.postproc(pb, cs, {"final_render": False})
            assert subset is None
            chunks.append(np.array(post_array, dtype=np.float64))
        fields = np.concatenate(chunks, axis=1)
        res = {"Z": out["Z"], "stop_iter": out["stop_iter"], "stop_reason": out["stop_reason"],
               "meta": json.dumps({"case": name, "posts": posts, "codes": list(codes),
                                   "px_snap": None, "floor_iter": 0,
                                   "reference": "GBillotey/Fractalshades v1.2.1, fastmath as shipped, "
                                                "postproc_dtype float64"})}
        for i, p in enumerate(posts):
            res[p] = fields[i]
        path = os.path.join(REPO, "tests", "golden", f"pp_{name}.npz")
        np.savez_compressed(path, **res)
        esc = out["stop_reason"][0] == 1
        print(name, posts, "pts", fields.shape[1], "escaped", int(esc.sum()),
              "size", os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv[1:])
