#!/bin/bash
# A/B of libkcrn variants on one box:  gpurun --timeout 400 -- 'bash profiles/run_ab.sh <tag> [shapes...]'
TAG=${1:-ab}; shift
SHAPES=${@:-rcv1 w8a}
O=gpurun_out; mkdir -p $O
( timeout 200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log )
tail -4 $O/${TAG}_pytest.log
{
for so in profiles/microbench/_build/libkcrn_*.so; do
  KCRN_LIBRARY=$PWD/$so timeout 120 python profiles/microbench/variant_time.py $SHAPES 2>&1 | grep -v "^$"
done
timeout 120 python profiles/microbench/variant_time.py $SHAPES 2>&1
KCRN_NO_GATHER_ORDER=1 timeout 120 python profiles/microbench/variant_time.py rcv1 2>&1 | sed 's/^default/no-gather-order/'
} > $O/${TAG}_variants.txt
cat $O/${TAG}_variants.txt
timeout 200 python bench.py --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c1-330 $O/${TAG}_bench.json
