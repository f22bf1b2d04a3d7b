#!/bin/bash
# run quick_bench (keys u32, variant 2) with every sweep library
for lib in tools/sweep/*.so; do
  echo "== $lib"
  OSB200_LIB=$PWD/$lib OSB_SKIP_PAIRS=1 OSB_VARIANTS=2 timeout 120 python tools/quick_bench.py 30 2>&1 | grep -E "variant=2 rank"
done
