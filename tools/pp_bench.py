# -*- coding: utf-8 -*-
""" End-to-end time of one frame: raw fields to the host (numba_cycle_call)
vs fused post-processing (postproc.frame_fields).  python tools/pp_bench.py config2 """
import json, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    w = bench.WORKLOADS[wl]
    from fractalshades_b200 import postproc as fpp
    f = bench.make_fractal(w)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    indep = f._calc_data["bench"]["cycle_indep_args"]
    st = f._calc_data["bench"]["state"]
    shapes = []
    c_pix = bench.frame_c_pix(f, shapes=shapes)
    n = c_pix.shape[0]
    from fractalshades_b200 import _native
    Z = _native.pinned_empty((len(st.codes[0]), n), st.complex_type)
    U = _native.pinned_empty((1, n), np.int32)
    sr = _native.pinned_empty((1, n), np.int8); si = _native.pinned_empty((1, n), np.int32)
    cp = _native.pinned_empty((n,), np.complex128); cp[:] = c_pix; c_pix = cp
    raw, fused = [], []
    for k in range(4):
        t0 = time.perf_counter()
        assert f.numba_cycle_call((c_pix, Z, U, sr, si), indep, tiles=shapes) == 0
        raw.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        out, stats = fpp.frame_fields(f, "bench", copy=False)
        fused.append(time.perf_counter() - t0)
    bytes_raw = Z.nbytes + U.nbytes + sr.nbytes + si.nbytes
    bytes_pp = sum(v.nbytes for v in out.values())
    print(json.dumps({"workload": wl, "raw_ms": round(1e3 * min(raw[1:]), 2),
                      "fused_pp_ms": round(1e3 * min(fused[1:]), 2),
                      "d2h_bytes_raw": bytes_raw, "d2h_bytes_pp": bytes_pp,
                      "kernel_ms": round(stats["kernel_ms"], 2)}))


if __name__ == "__main__":
    main()
