import sys, os, time
sys.path.insert(0, "/root/repo")
import bench
from fractalshades_b200 import settings
w = bench.WORKLOADS["config2"]
f = bench.make_fractal(w)
f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
def reset():
    d = os.path.join(f.directory, "data")
    for n in os.listdir(d):
        if n.startswith("bench_") and n.endswith(".arr") or n == "bench.report":
            os.unlink(os.path.join(d, n))
    f._calc_data["bench"]["need_new_mmap"] = True
for nt in (8, 8, 12, 16, 24, 32):
    settings.io_threads = nt
    reset()
    t0 = time.perf_counter(); f.calc_raw("bench"); t1 = time.perf_counter()
    print("threads", nt, "calc_raw ms %.1f" % ((t1 - t0) * 1e3), "cpus", os.cpu_count())
settings.io_threads = 16
t0 = time.perf_counter(); f._calc_data["bench"]["need_new_mmap"] = False
rep = f.get_report_memmap("bench", mode="r+"); rep[:, 3] = 0; rep.flush()
f.calc_raw("bench"); t1 = time.perf_counter()
print("rewrite of existing files ms %.1f" % ((t1 - t0) * 1e3))
import subprocess; print(subprocess.run("df -h /tmp | tail -1; mount | grep -E ' / | /tmp ' | head -3", shell=True, capture_output=True, text=True).stdout)
