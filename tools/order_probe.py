# -*- coding: utf-8 -*-
"""
Probe: how much does the shape of a warp's 32-pixel footprint matter?
The kernels take any point list, so the footprint is changed here by reordering
c_pix on the host (P columns x Q rows patches inside each tile) and timing the
device-resident launch.   python tools/order_probe.py --workload config3
"""
import argparse, os, sys, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench


def ordered(f, P, Q):
    out = []
    for cs in f.chunk_slices():
        t = f.chunk_pixel_pos(cs, False, None)          # (h, w) complex? or flat
        ix, ixx, iy, iyy = cs
        w_, h_ = ixx - ix, iyy - iy
        a = np.ravel(t).reshape(h_, w_) if t.shape != (h_, w_) else t
        if a.shape != (h_, w_):
            a = np.ravel(t).reshape(w_, h_)
        hh, ww = a.shape
        if hh % Q or ww % P:
            out.append(np.ravel(a)); continue
        out.append(a.reshape(hh // Q, Q, ww // P, P).transpose(0, 2, 1, 3).ravel())
    return np.ascontiguousarray(np.concatenate(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--shapes", default="32x1,16x2,8x4,4x8,2x16")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    from fractalshades_b200 import _native
    lib = _native.cuda_lib()
    f = bench.make_fractal(w)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    indep = f._calc_data["bench"]["cycle_indep_args"]
    frame = indep[1]
    state = f._calc_data["bench"]["state"]
    n_Z, n_U = len(state.codes[0]), len(state.codes[1])
    zdt = np.dtype(state.complex_type)
    ref = None
    for shp in args.shapes.split(","):
        P, Q = (int(v) for v in shp.split("x"))
        c = ordered(f, P, Q)
        npts = c.shape[0]
        d_c = lib.fsb_dev_alloc(npts * 16); d_Z = lib.fsb_dev_alloc(n_Z * npts * zdt.itemsize)
        d_U = lib.fsb_dev_alloc(max(n_U, 1) * npts * 4); d_sr = lib.fsb_dev_alloc(npts)
        d_si = lib.fsb_dev_alloc(npts * 4)
        _native.check(lib, lib.fsb_memcpy_h2d(d_c, _native.ptr(c), npts * 16))
        st = _native.FsbStats()
        ms = []
        for k in range(5):
            lib.fsb_flush_l2()
            _native.check(lib, lib.fsb_frame_run_device(frame.ptr, npts, d_c, d_Z, d_U, d_sr, d_si, st))
            ms.append(st.kernel_ms)
        if ref is None:
            ref = st.sum_stop_iter
        assert st.sum_stop_iter == ref, (st.sum_stop_iter, ref)
        print(json.dumps({"workload": args.workload, "footprint": shp, "kernel_ms": round(float(np.median(ms[2:])), 3),
                          "n_iter_exec": st.n_iter_exec, "n_bla": st.n_bla_steps}), flush=True)
        for p in (d_c, d_Z, d_U, d_sr, d_si):
            lib.fsb_dev_free(p)


if __name__ == "__main__":
    main()
