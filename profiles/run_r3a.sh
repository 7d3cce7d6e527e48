#!/bin/bash
# recurrence passes with the deferred sums on one warp, the control block behind the gathers, speculative gather sources
O=gpurun_out; mkdir -p $O; TAG=${1:-r3a}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_pytest.log; tail -5 $O/${TAG}_pytest.log
{
for g in 1,2 2,2 auto; do
  if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
  echo "=== KCRN_GATHER=$g"
  timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos graph|Error|error"
done
} > $O/${TAG}_variants.txt 2>&1
cat $O/${TAG}_variants.txt
