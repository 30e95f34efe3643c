# -*- coding: utf-8 -*-
""" Create the device frame of a bench workload twice and print the setup timings. """
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config3"]
f = bench.make_fractal(w, 256)
for i in range(2):
    t0 = time.time()
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    fr = f._calc_data["bench"]["cycle_indep_args"][1]
    print(i, "calc_std_div s", round(time.time() - t0, 3), fr.setup_ms())
