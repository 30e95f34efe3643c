# -*- coding: utf-8 -*-
""" Fast-path Xrange kernel vs the pure Xrange kernel (FSB200_PURE_XR=1) on a
sizeable config-3 frame: the strict builds must agree bit for bit. """
import os, sys, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np

def run(out, nx, workload, strict):
    import bench
    from fractalshades_b200 import settings
    settings.strict_ieee = bool(strict)
    w = bench.WORKLOADS[workload]
    f = bench.make_fractal(w, nx)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    indep = f._calc_data["bench"]["cycle_indep_args"]
    c_pix = bench.frame_c_pix(f)
    n = c_pix.shape[0]
    st = f._calc_data["bench"]["state"]
    Z = np.zeros((len(st.codes[0]), n), st.complex_type); U = np.zeros((1, n), np.int32)
    sr = -np.ones((1, n), np.int8); si = np.zeros((1, n), np.int32)
    assert f.numba_cycle_call((c_pix, Z, U, sr, si), indep) == 0
    from fractalshades_b200 import Fractal
    print(os.environ.get("FSB200_PURE_XR", "0"), "kernel ms", Fractal._last_stats["kernel_ms"], "sum", int(si.sum(dtype=np.int64)))
    np.savez(out, Z=Z, U=U, sr=sr, si=si)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        run(sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5]))
        sys.exit(0)
    nx = int(os.environ.get("NX", "1280"))
    for workload in os.environ.get("WORKLOADS", "config3,config4").split(","):
        for strict in (1, 0):
            d = tempfile.mkdtemp()
            outs = []
            for pure in ("0", "1"):
                o = os.path.join(d, f"o{pure}.npz")
                env = dict(os.environ, FSB200_PURE_XR=pure)
                subprocess.check_call([sys.executable, os.path.abspath(__file__), "--worker", o, str(nx), workload, str(strict)], env=env)
                outs.append(np.load(o))
            a, b = outs
            same_i = np.mean(a["si"] == b["si"]); same_r = np.mean(a["sr"] == b["sr"]); same_u = np.mean(a["U"] == b["U"])
            zb = np.mean((a["Z"] == b["Z"]) | (np.isnan(a["Z"]) & np.isnan(b["Z"])))
            print(f"{workload} strict={strict}: fast vs pure Xrange: stop_iter {same_i*100:.5f}% reason {same_r*100:.5f}% U {same_u*100:.5f}% Z bits {zb*100:.5f}%", flush=True)
