#!/bin/bash
# Lanczos passes: tile values requested behind the gathers (KCRN_LATE_VALUES 1 = default build, 0 = at entry, 2 = in front of the multiply); alternating plain passes
O=gpurun_out; mkdir -p $O; TAG=${1:-r2z}
{
for lib in default lv0 lv2 default lv0 lv2; do
  if [ $lib = default ]; then unset KCRN_LIBRARY; else export KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_$lib.so; fi
  for g in 1,2 2,2 auto; do
    if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
    echo "=== lib=$lib KCRN_GATHER=$g"
    timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos graph|alternating|Error|error"
  done
done
} > $O/${TAG}_variants_late_values.txt 2>&1
cat $O/${TAG}_variants_late_values.txt
