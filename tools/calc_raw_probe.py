# -*- coding: utf-8 -*-
""" Time the public API end to end: zoom -> calc_std_div -> calc_raw (memmaps). """
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config2"]
f = bench.make_fractal(w)
t0 = time.time(); f.calc_std_div(calc_name="bench", subset=None, **w["calc"]); t1 = time.time()
print("calc_std_div (orbit + tables) s", round(t1 - t0, 3))
for rep in range(2):
    f.clean_up("bench"); f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    t0 = time.time()
    if rep == 1:
        pr = cProfile.Profile(); pr.enable()
    f.calc_raw("bench")
    if rep == 1:
        pr.disable()
    t1 = time.time()
    print("calc_raw s", round(t1 - t0, 3), f.last_stats)
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
