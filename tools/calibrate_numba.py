# -*- coding: utf-8 -*-
"""
calibrate_numba.py -- time the REFERENCE's numba `calc_raw` beside the oracle
port (oracle/fs_oracle.cpp, the CPU arm of bench.py) on the same pixels, in the
build container (the reference is a Python package: it cannot travel to the GPU
box, so bench.py's `cpu_baseline` / `--impl reference` time the port there and
carry this ratio).  BASELINE.md section 3: the timed span is `f.calc_raw(calc)`
(all tiles through compute_rawdata_dev incl. the memmap writes) on an instance
whose numba kernels are already compiled, host threads = os.cpu_count().

    python tools/calibrate_numba.py [--nx 1280] [--workloads config2,config3]
writes profiles/numba_calibration.json

TEST INFRASTRUCTURE: imports /root/reference through tools/ref_harness.py.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1280)
    ap.add_argument("--workloads", default="config2,config3")
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "numba_calibration.json"))
    args = ap.parse_args()
    import bench
    import ref_harness as rh
    import oracle_lib as ol
    import parity_common as pc
    res = {}
    if os.path.exists(args.out):
        res = json.load(open(args.out))
    for wname in args.workloads.split(","):
        w = bench.WORKLOADS[wname]
        case = dict(kind=w["kind"], x=w["x"], y=w["y"], dx=w["dx"], nx=args.nx,
                    xy_ratio=w["xy_ratio"], theta_deg=w.get("theta_deg", 0.),
                    precision=w.get("precision"), calc=w["calc"], init=w.get("init", {}),
                    skew=w.get("skew", {}))
        fs = rh.load_reference()
        cores = os.cpu_count()
        d = tempfile.mkdtemp(prefix="fs_cal_")
        t0 = time.time()
        f = rh.make_fractal(case, d)          # orbit, dZndc path, BLA tree (reference code)
        t_setup = time.time() - t0
        t0 = time.time()
        f.calc_raw("c")                       # first call: numba compilation + run
        t_first = time.time() - t0
        f.clean_up("c")
        f.calc_std_div(calc_name="c", subset=None, **w["calc"])
        t0 = time.time()
        f.calc_raw("c")                       # warm: the timed span of BASELINE.md section 3
        t_numba = time.time() - t0
        si_ref = np.array(f.get_data_memmap("c", "stop_iter", mode="r"))
        iters = int(si_ref.sum(dtype=np.int64))
        # the port on the same pixels, tables built by the product's host code
        fb = bench.make_fractal(w, args.nx)
        spec = bench.bind_spec(fb, w)
        t, t_tables = bench.oracle_frame(w, fb, spec)
        c_pix = bench.frame_c_pix(fb)
        bench.oracle_run(w, fb, t, c_pix[:20000], cores)       # warm
        t0 = time.time()
        si_port = bench.oracle_run(w, fb, t, c_pix, cores)
        t_port = time.time() - t0
        same = float(np.mean(si_port.ravel() == si_ref.ravel()))
        res[wname] = {
            "nx": args.nx, "npts": int(c_pix.shape[0]), "cores": cores,
            "numba_calc_raw_s": t_numba, "numba_first_call_s": t_first,
            "numba_gpix_iter_s": iters / t_numba / 1e9,
            "port_s": t_port, "port_gpix_iter_s": int(si_port.sum(dtype=np.int64)) / t_port / 1e9,
            "port_over_numba": t_numba / t_port,
            "stop_iter_equal_fraction": same,
            "reference_setup_s": t_setup, "port_tables_s": t_tables,
            "where": "build container (no GPU); numba %s" % __import__("numba").__version__,
            "note": "numba = the reference's own calc_raw (fastmath, all tiles, memmap writes "
                    "included), warm instance; port = oracle/fs_oracle.cpp with the same threads",
        }
        print(wname, json.dumps(res[wname]))
        shutil.rmtree(d, ignore_errors=True)
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
