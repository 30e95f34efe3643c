#!/bin/bash
# A/B kernel timing: runs bench.py for each library under ab/ (FSB200_LIB
# override) on the given workloads; prints "lib workload kernel_ms e2e_ms".
# usage: tools/ab_bench.sh "config2 config3" ab/base.so ab/v1.so ...
wl="$1"; shift
for w in $wl; do
  for lib in "$@"; do
    out=$(FSB200_LIB=$PWD/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
    python - "$lib" "$w" <<PY
import json,sys
d=json.loads('''$out''')
print(sys.argv[1], sys.argv[2], "kernel_ms=%.2f e2e_ms=%.2f frac=%.3f exec=%d bla=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["n_iter_exec"], d["roofline"]["n_bla_steps"]))
PY
  done
done
