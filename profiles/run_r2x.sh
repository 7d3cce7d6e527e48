#!/bin/bash
# 3 CTAs of 14 items per SM (80 registers) against the default 4 x 11 (64 registers), same box
O=gpurun_out; mkdir -p $O; TAG=${1:-r2x}
{
for lib in default mb3 default mb3; do
  if [ $lib = default ]; then unset KCRN_LIBRARY; else export KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_$lib.so; fi
  for g in 1,2 2,2; do
    echo "=== lib=$lib KCRN_GATHER=$g"
    KCRN_VERBOSE=1 KCRN_GATHER=$g timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos graph|tiles of|Error|error"
  done
done
} > $O/${TAG}_variants.txt 2>&1
cat $O/${TAG}_variants.txt
