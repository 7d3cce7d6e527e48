#!/bin/bash
# gather variant 3 (asynchronous gathers in two halves): bit equality + timing
O=gpurun_out; mkdir -p $O; TAG=${1:-r3b}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or switches or hvp or lanczos" > $O/${TAG}_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_pytest.log; tail -5 $O/${TAG}_pytest.log
{
for g in 1,2 2,2 3,3 1,3 3,2 auto; do
  if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
  echo "=== KCRN_GATHER=$g"
  KCRN_VERBOSE=1 timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos graph|alternating|gather variant|Error|error"
done
} > $O/${TAG}_variants.txt 2>&1
grep -v "gather variant" $O/${TAG}_variants.txt
