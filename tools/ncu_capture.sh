#!/bin/bash
# ncu evidence of the pixel kernels (run under gpurun, one GPU):
#   gpurun_out/<tag>_config{2,3,4}.ncu-rep  --set full capture of the first (full-frame, warm-up) launch
#   gpurun_out/<tag>_launches.csv           launch list of the default bench command
# usage: tools/ncu_capture.sh <tag>
tag=${1:-cap}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in config2 config3 config4; do
  ncu --set full --clock-control none --import-source on -k regex:k_perturb -s 0 -c 1 -f \
      -o gpurun_out/${tag}_$w python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline \
      > gpurun_out/${tag}_$w.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/${tag}_launches.log 2>&1
ls -la gpurun_out | grep $tag
