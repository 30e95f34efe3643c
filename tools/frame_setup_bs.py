import sys, os, time
sys.path.insert(0, "/root/repo")
import bench
w = bench.WORKLOADS["config4"]
for host in ("1", "0"):
    os.environ["FSB200_HOST_DZNDC"] = host
    f = bench.make_fractal(w)
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    fr = f._calc_data["bench"]["cycle_indep_args"][1]
    f.clean_up("bench")
    f.calc_std_div(calc_name="bench", subset=None, **w["calc"])
    fr = f._calc_data["bench"]["cycle_indep_args"][1]
    print("FSB200_HOST_DZNDC =", host, fr.setup_ms(), "L =", len(fr.tables["Zn_path"]))
