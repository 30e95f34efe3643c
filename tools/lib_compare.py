# -*- coding: utf-8 -*-
""" Render one frame with two builds of libfsb200 (FSB200_LIB override) and
compare the outputs bit for bit.
    python tools/lib_compare.py libA.so libB.so [workload] [nx] """
import os, sys, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
import numpy as np


def worker(out, workload, nx):
    import bench
    w = bench.WORKLOADS[workload]
    f = bench.make_fractal(w, nx)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    indep = f._calc_data["bench"]["cycle_indep_args"]
    shapes = []
    c_pix = bench.frame_c_pix(f, shapes=shapes)
    n = c_pix.shape[0]
    st = f._calc_data["bench"]["state"]
    Z = np.zeros((len(st.codes[0]), n), st.complex_type); U = np.zeros((1, n), np.int32)
    sr = -np.ones((1, n), np.int8); si = np.zeros((1, n), np.int32)
    assert f.numba_cycle_call((c_pix, Z, U, sr, si), indep, tiles=shapes) == 0
    from fractalshades_b200 import Fractal
    print(os.path.basename(os.environ["FSB200_LIB"]), "kernel ms", round(Fractal._last_stats["kernel_ms"], 2),
          {k: Fractal._last_stats[k] for k in ("n_iter_exec", "n_bla_steps", "n_rebase", "sum_stop_iter")})
    np.savez(out, Z=Z, U=U, sr=sr, si=si)


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(sys.argv[2], sys.argv[3], int(sys.argv[4])); sys.exit(0)
    la, lb = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else "config3"
    nx = sys.argv[4] if len(sys.argv) > 4 else "1280"
    d = tempfile.mkdtemp(); outs = []
    for k, lib in enumerate((la, lb)):
        o = os.path.join(d, f"o{k}.npz")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--worker", o, workload, nx],
                              env=dict(os.environ, FSB200_LIB=os.path.abspath(lib)))
        outs.append(np.load(o))
    a, b = outs
    same = lambda x, y: float(np.mean(x == y)) * 100
    zb = float(np.mean(a["Z"].view(np.uint64) == b["Z"].view(np.uint64))) * 100
    print(f"{workload} nx={nx}: stop_iter {same(a['si'], b['si']):.5f}% reason {same(a['sr'], b['sr']):.5f}% "
          f"U {same(a['U'], b['U']):.5f}% Z bits {zb:.5f}%")
