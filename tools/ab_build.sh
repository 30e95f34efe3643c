#!/bin/bash
# Build A/B variants of the CUDA library into ab/ (git-ignored, travels with gpurun).
# usage: tools/ab_build.sh name1 "flags1" name2 "flags2" ...   (run in parallel)
cd "$(dirname "$0")/.."
mkdir -p ab
pids=()
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -shared \
      -cudart static -ccbin /usr/bin/g++ $flags -o ab/$name.so fractalshades_b200/csrc/fsb200.cu \
      > ab/$name.log 2>&1 && echo "built ab/$name.so" || { echo "FAILED ab/$name"; tail -5 ab/$name.log; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
