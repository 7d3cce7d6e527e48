#!/bin/bash
# A/B of libkcrn variants on one box:  gpurun --timeout 400 -- 'bash profiles/run_ab.sh <tag> [shapes...]'
# (variants: profiles/microbench/build_variants.sh -> profiles/microbench/_build/libkcrn_<name>.so)
TAG=${1:-ab}; shift
SHAPES=${@:-rcv1 w8a}
O=gpurun_out; mkdir -p $O
( timeout 200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log )
tail -3 $O/${TAG}_pytest.log
for so in profiles/microbench/_build/libkcrn_*.so; do
  case $so in *base*) continue;; esac
  ( KCRN_LIBRARY=$PWD/$so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${TAG}_pytest_$(basename $so .so).log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_$(basename $so .so).log )
  echo $so; tail -3 $O/${TAG}_pytest_$(basename $so .so).log
done
{
for so in profiles/microbench/_build/libkcrn_*.so; do
  KCRN_LIBRARY=$PWD/$so timeout 120 python profiles/microbench/variant_time.py $SHAPES 2>&1 | grep -v "^$"
done
} > $O/${TAG}_variants.txt
cat $O/${TAG}_variants.txt
