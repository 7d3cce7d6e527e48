#!/bin/bash
# knock-out timing of the Lanczos-specific parts of the recurrence passes (-DKCRN_KO=bits builds, wrong results by design)
O=gpurun_out; mkdir -p $O; TAG=${1:-r2z}
rm -f profiles/microbench/_build/libkcrn_lv*.so profiles/microbench/_build/libkcrn_mb*.so
{
for lib in default ko1 ko2 ko4 ko8 ko15 default; do
  if [ $lib = default ]; then unset KCRN_LIBRARY; else export KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_$lib.so; fi
  for g in 1,2 2,2; do
    export KCRN_GATHER=$g
    echo "=== lib=$lib KCRN_GATHER=$g"
    timeout 100 python profiles/microbench/variant_time.py rcv1 2>&1 | grep -E "Lanczos graph|Error|error"
  done
done
} > $O/${TAG}_knockout.txt 2>&1
cat $O/${TAG}_knockout.txt
