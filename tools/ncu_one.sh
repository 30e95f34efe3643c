#!/bin/bash
# ncu --set full capture of the first full-frame pixel-kernel launch of one workload.
# usage: tools/ncu_one.sh <tag> <workload> [lib.so]
tag=$1; w=$2; lib=$3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
[ -n "$lib" ] && export FSB200_LIB=$PWD/$lib
ncu --set full --clock-control none --import-source on -k regex:k_perturb -s 0 -c 1 -f \
    -o gpurun_out/${tag}_$w python bench.py --workload $w --also none --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/${tag}_$w.log 2>&1
tail -2 gpurun_out/${tag}_$w.log | cut -c1-300
