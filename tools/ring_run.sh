#!/bin/bash
# time the sweep libraries with quick_bench: ring_* with the persistent TMA-ring kernel (variant 1), anything else with the
# default kernel; OSB_SKIP_PAIRS= (empty) also times pairs and u64
for lib in tools/sweep/*.so; do
  v=2; case "$lib" in *ring*) v=1;; esac
  echo "== $lib (variant $v)"
  OSB200_LIB=$PWD/$lib OSB_VARIANTS=$v timeout 240 python tools/quick_bench.py ${1:-30} 2>&1 | grep -E "variant=$v rank|pairs|u64|Error|error|assert" | grep -v REFERENCE
done
echo "== product lib"
OSB_VARIANTS=2 timeout 240 python tools/quick_bench.py ${1:-30} 2>&1 | grep -E "variant=2 rank|pairs|u64|Error|error|assert"
