#!/bin/bash
# run quick_bench (keys u32, default variant) with every sweep library; prints the whole-sort median and one-pass time
for lib in tools/sweep/*.so; do
  echo "== $lib"
  OSB200_LIB=$PWD/$lib OSB_SKIP_PAIRS=${OSB_SKIP_PAIRS-1} OSB_VARIANTS=2 timeout 180 python tools/quick_bench.py ${1:-30} 2>&1 | grep -E "variant=2|pairs|u64|Error|error|assert" | grep -v REFERENCE
done
