#!/bin/bash
# Build the library with alternative compile-time parameters into tools/sweep/ (development sweeps; see tools/sweep_run.sh).
# each config: "WARPS K MINB LOOK STEP EXP [ABL]"   (EXP/ABL: OSB_EXP experiment / OSB_ABL ablation bits, see osb_kernels.cu)
set -e
cd "$(dirname "$0")/../gpusorting_b200/csrc"
mkdir -p ../../tools/sweep
rm -f ../../tools/sweep/*.so
CONFIGS=${CONFIGS:-"16 32 2 16 8 0;16 32 2 32 8 0;16 32 2 16 8 4"}
IFS=';' read -ra CFGS <<< "$CONFIGS"
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  abl=${7:-0}
  out=../../tools/sweep/libosb_W$1_K$2_B$3_L$4_S$5_E$6_A$abl.so
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -DOSB200_BUILDING -DOSB_WIDE_WARPS=$1 -DOSB_WIDE_K=$2 -DOSB_WIDE_MINB=$3 -DOSB_LOOK=$4 -DOSB_STEP=$5 -DOSB_EXP=$6 -DOSB_ABL=$abl $EXTRA_DEFS -Xptxas -v \
       -shared -o $out osb_kernels.cu osb_host.cu osb_sharded.cu -lnccl 2> ../../tools/sweep/build_W$1_K$2_B$3_L$4_S$5_E$6_A$abl.log &
done
wait
ls -la ../../tools/sweep/*.so
