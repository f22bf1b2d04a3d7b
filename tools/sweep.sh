#!/bin/bash
# Build the library with alternative lookback window parameters into tools/sweep/ (development sweeps).
set -e
cd "$(dirname "$0")/../gpusorting_b200/csrc"
for cfg in "8 4" "8 8" "24 8" "32 8" "16 4" "16 16"; do
  set -- $cfg
  out=../../tools/sweep/libosb_L$1_S$2.so
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -DOSB200_BUILDING -DOSB_LOOK=$1 -DOSB_STEP=$2 -shared -o $out osb_kernels.cu osb_host.cu osb_sharded.cu -lnccl &
done
wait
ls -la ../../tools/sweep/
