#!/bin/bash
# Build the library with alternative geometry / lookback parameters into tools/sweep/ (development sweeps).
# each config: "WARPS K MINB LOOK STEP"
set -e
cd "$(dirname "$0")/../gpusorting_b200/csrc"
rm -f ../../tools/sweep/*.so
for cfg in "8 32 4 16 8" "8 32 4 8 4" "10 32 3 8 4" "10 32 3 16 8"; do
  set -- $cfg
  out=../../tools/sweep/libosb_W$1_K$2_B$3_L$4_S$5.so
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -DOSB200_BUILDING -DOSB_WIDE_WARPS=$1 -DOSB_WIDE_K=$2 -DOSB_WIDE_MINB=$3 -DOSB_LOOK=$4 -DOSB_STEP=$5 -Xptxas -v \
       -shared -o $out osb_kernels.cu osb_host.cu osb_sharded.cu -lnccl 2> ../../tools/sweep/build_W$1_K$2_B$3_L$4_S$5.log &
done
wait
ls -la ../../tools/sweep/*.so
