#!/bin/bash
O=gpurun_out; mkdir -p $O
{
for e in "" "KCRN_NO_SPEC=1"; do
  echo "=== $e"
  env $e KCRN_VERBOSE=1 timeout 200 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos|sweeps"
done
} > $O/r2g_ab.txt 2>&1
cat $O/r2g_ab.txt
