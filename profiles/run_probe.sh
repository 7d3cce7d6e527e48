#!/bin/bash
O=gpurun_out; mkdir -p $O
{
for lib in "" "profiles/microbench/_build/libkcrn_noscatter.so"; do
  for g in "1,2" "1"; do
  echo "=== lib=$lib KCRN_GATHER=$g"
  if [ -n "$lib" ]; then export KCRN_LIBRARY=$PWD/$lib; else unset KCRN_LIBRARY; fi
  KCRN_GATHER=$g timeout 200 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos|sweeps"
  done
done
} > $O/r2p_noscatter.txt 2>&1
cat $O/r2p_noscatter.txt
