#!/bin/bash
# Tile positions from the kernels' parameter bank (no descriptor round trip before the stream): parity + same-box A/B
TAG=${1:-r2v}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; RC=$?
echo "exit $RC" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
{
for lib in default nopdesc default nopdesc; do
  if [ $lib = default ]; then unset KCRN_LIBRARY; else export KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_$lib.so; fi
  for g in 1,2 auto; do
    if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
    echo "=== lib=$lib KCRN_GATHER=$g"
    timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos graph|Error|error"
  done
done
} > $O/${TAG}_variants.txt 2>&1
unset KCRN_GATHER KCRN_LIBRARY
cat $O/${TAG}_variants.txt
