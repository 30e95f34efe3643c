# -*- coding: utf-8 -*-
""" Summarise an .ncu-rep (raw + source pages) into a short text report. """
import csv, subprocess, sys, io
from collections import Counter

def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return [r for r in csv.reader(io.StringIO(out)) if len(r) > 10]

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__cycles_active.avg"]

def main(rep, top=25):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    print("== kernel:", d.get("Kernel Name", ("", ""))[1][:120])
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][1]:>18s} {d[k][0]}")
    stalls = {k: float(v[1]) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    print("== warp stall reasons (avg warps stalled per issue-active cycle):")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
        print(f"   {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
    rows = page(rep, "source")
    hdr = rows[0]
    iS, iN, iI, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = rows[1:]
    tot = sum(int(r[iI] or 0) for r in data)
    c = Counter()
    for r in data:
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        c[op.split(".")[0]] += int(r[iI] or 0)
    print(f"== executed warp instructions: {tot}  ({len(data)} SASS lines); opcode mix:")
    print("   " + "  ".join(f"{op} {n / tot * 100:.1f}%" for op, n in c.most_common(22)))
    fp64 = sum(n for op, n in c.items() if op in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
    print(f"   FP64 instructions (DADD+DMUL+DFMA+DSETP): {fp64 / tot * 100:.1f}% of issued")

if __name__ == "__main__":
    main(sys.argv[1])
