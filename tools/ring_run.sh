#!/bin/bash
# time the sweep libraries: ring_* with the persistent TMA-ring kernel (variant 1), anything else with the default kernel
for lib in tools/sweep/*.so; do
  v=2; case "$lib" in *ring*) v=1;; esac
  echo "== $lib (variant $v)"
  OSB200_LIB=$PWD/$lib OSB_SKIP_PAIRS=1 OSB_VARIANTS=$v timeout 180 python tools/quick_bench.py ${1:-30} 2>&1 | grep -E "variant=$v|Error|error|assert" | grep -v REFERENCE
done
echo "== product lib"
OSB_SKIP_PAIRS=1 OSB_VARIANTS=2 timeout 180 python tools/quick_bench.py ${1:-30} 2>&1 | grep -E "variant=2|Error|error|assert"
