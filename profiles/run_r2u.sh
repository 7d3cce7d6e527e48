#!/bin/bash
# Split Lanczos passes (KCRN_SPLIT) and the bulk-word variant with early gathers: parity subset + same-box A/B
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or gather or fused or cubic or lanczos" > $O/${TAG}_pytest.log 2>&1; RC=$?
echo "exit $RC" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
{
for sp in 0 1; do
  for g in 1,2 1,3 0 auto; do
    if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
    echo "=== KCRN_SPLIT=$sp KCRN_GATHER=$g"
    KCRN_SPLIT=$sp timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos graph|another|Error|error"
  done
done
} > $O/${TAG}_variants.txt 2>&1
unset KCRN_GATHER
cat $O/${TAG}_variants.txt
for sp in 0 1; do
KCRN_SPLIT=$sp timeout 250 python bench.py --no-cpu-baseline > $O/${TAG}_bench_split$sp.json 2> $O/${TAG}_bench_split$sp.err
echo "bench exit $?" >> $O/${TAG}_bench_split$sp.err
python - <<PY
import json
j=json.loads(open("$O/${TAG}_bench_split$sp.json").read().strip().splitlines()[-1])
print("split$sp", j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["hvp"], j.get("device_loop"))
PY
done
