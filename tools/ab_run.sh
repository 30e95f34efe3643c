#!/bin/bash
# A/B helper for kernel experiments on the GPU box: benches the in-tree library on
# configs 2-4, then every ab/*.so on the workloads given as arguments (default config3).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
for w in config2 config3 config4; do
  echo "## default $w" >> gpurun_out/ab.jsonl
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/ab.jsonl 2>gpurun_out/ab_err.log
done
for v in ab/*.so; do
  [ -e "$v" ] || continue
  for w in ${@:-config3}; do
    echo "## $v $w" >> gpurun_out/ab.jsonl
    FSB200_LIB=$PWD/$v python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/ab.jsonl 2>>gpurun_out/ab_err.log
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/ab.jsonl'):
    if l.startswith('##'): print(l.strip()); continue
    try:
        d=json.loads(l); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']), round(r['frac'],4), round(d['e2e']['ms_per_step'],2), r['n_iter_exec'], r['n_bla_steps'], r['n_rebase'], r['n_iter_fast'])
    except Exception as e: print('ERR', l[:200])
PY
