#!/bin/bash
# A/B helper for kernel experiments on the GPU box: benches the in-tree library (and,
# with FSB200_KERNEL_V1=1, the general round-1 kernel) then every ab/*.so, on the
# workloads given as arguments (default: config2 config3).  One line per run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
: > gpurun_out/ab_err.log
WL=${@:-config2 config3}
B="python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline"
for w in $WL; do
  echo "## v1-kernel $w" >> gpurun_out/ab.jsonl
  FSB200_KERNEL_V1=1 $B --workload $w >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
  echo "## in-tree $w" >> gpurun_out/ab.jsonl
  $B --workload $w >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
done
for v in ab/*.so; do
  [ -e "$v" ] || continue
  for w in $WL; do
    echo "## $v $w" >> gpurun_out/ab.jsonl
    FSB200_LIB=$PWD/$v $B --workload $w >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/ab.jsonl'):
    if l.startswith('##'): print(l.strip()); continue
    try:
        d=json.loads(l); r=d['roofline']; print("   ms %.3f  Gpix-it/s %d  frac %.4f  e2e_ms %.2f  exec %d bla %d reb %d fast %d" % (d['ms_per_step'], d['value'], r['frac'], d['e2e']['ms_per_step'], r['n_iter_exec'], r['n_bla_steps'], r['n_rebase'], r['n_iter_fast']))
    except Exception as e: print('ERR', l[:300])
PY
tail -5 gpurun_out/ab_err.log
