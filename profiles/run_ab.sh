#!/bin/bash
O=gpurun_out; mkdir -p $O
{
for e in "" "KCRN_GATHER=1" "KCRN_GATHER=1,2"; do
  echo "=== $e"
  env $e KCRN_VERBOSE=1 timeout 200 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos|sweeps"
done
} > $O/${1:-r2n}_ab.txt 2>&1
grep -E "===|default" $O/${1:-r2n}_ab.txt
