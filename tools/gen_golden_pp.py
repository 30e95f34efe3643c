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

PP_CASES = ["std_M2_cfg1", "p_M2_E20", "p_M2_deep250", "p_BS_f1_E30_skew", "std_BS_f1",
            # with a stored back-shifted orbit (field lines start from zn_orbit / xn_orbit)
            "std_M2_seahorse_orbit", "p_M2_divref_orbit", "p_BS_f2_E12"]
# Fieldlines_pp(**FL) and the light sources of the shading fixture
FL = dict(n_iter=4, swirl=0.3, endpoint_k=0.5)
LIGHTS = [dict(k_diffuse=1.8, k_specular=15., shininess=500., polar_angle=50., azimuth_angle=20.),
          dict(k_diffuse=0.4, k_specular=0., shininess=400., polar_angle=-65., azimuth_angle=35.)]
MAX_SLOPE = 60.


def main(argv):
    import ref_harness as rh
    from cases import CASES
    fs = rh.load_reference()
    fs.settings.postproc_dtype = "float64"
    from fractalshades.postproc import (Postproc_batch, Continuous_iter_pp, DEM_pp,
                                        DEM_normal_pp, Fieldlines_pp)
    from fractalshades.colors.layers import Blinn_lighting
    names = argv or PP_CASES
    for name in names:
        case = CASES[name]
        out = rh.run_case(case, keep_tables=False)
        f = out["fractal"]
        codes = f._calc_data["c"]["saved_codes"][0]
        have_deriv = ("dzndc" in codes) or ("dxnda" in codes)
        pb = Postproc_batch(f, "c")
        pb.add_postproc("cont_iter", Continuous_iter_pp())
        pb.add_postproc("fieldlines", Fieldlines_pp(**FL))
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
                                   "px_snap": None, "floor_iter": 0, "fieldlines": FL,
                                   "lights": LIGHTS, "max_slope": MAX_SLOPE,
                                   "reference": "GBillotey/Fractalshades v1.2.1, fastmath as shipped, "
                                                "postproc_dtype float64"})}
        for i, p in enumerate(posts):
            res[p] = fields[i]
        res["c_pix"] = out["c_pix"]
        if have_deriv:
            # Color_layer.apply_shade (colors/layers.py:493-507) + the reference's own
            # Blinn_lighting.partial_shade (:865-903) on a white XYZ image: with
            # (k_diffuse, k_specular) = (1, 0) it returns the Lambert coefficient, with
            # (0, 1) the specular one
            normal = np.array([fields[posts.index("normal_x")],
                               fields[posts.index("normal_y")]]).reshape(2, 1, -1)
            complex_n = np.empty(shape=normal.shape[1:], dtype=np.complex64)
            coeff = np.sin(MAX_SLOPE * np.pi / 180)
            complex_n.real = normal[0, :, :] * coeff
            complex_n.imag = normal[1, :, :] * coeff
            XYZ = np.ones(complex_n.shape + (3,))
            rows = []
            with np.errstate(all="ignore"):
                for ls in LIGHTS:
                    lt = Blinn_lighting(0.2, (1., 1., 1.))
                    lt.add_light_source(**dict(ls, k_diffuse=1., k_specular=0.))
                    rows.append(lt.partial_shade(lt.light_sources[0], XYZ, complex_n)[0, :, 0])
                    lt = Blinn_lighting(0.2, (1., 1., 1.))
                    lt.add_light_source(**dict(ls, k_diffuse=0., k_specular=1.))
                    spec = lt.partial_shade(lt.light_sources[0], XYZ, complex_n)[0, :, 0]
                    rows.append(spec if ls["k_specular"] != 0. else np.zeros_like(spec))
            res["shade"] = np.array(rows, dtype=np.float64)
        path = os.path.join(REPO, "tests", "golden", f"pp_{name}.npz")
        np.savez_compressed(path, **res)
        esc = out["stop_reason"][0] == 1
        print(name, posts, "pts", fields.shape[1], "escaped", int(esc.sum()),
              "size", os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv[1:])
