#!/bin/bash
# First GPU call of the next round (one B200, about 4 minutes):  gpurun --timeout 420 -- 'bash profiles/run_next_round.sh r2a'
#   1. the GPU suite including the checks that could not run at the end of round 1 (experimental gather variant 3 in a
#      child process; the one-GPU emulation of the row-sharded code path, a non-strict xfail until it has passed once)
#   2. gather variants side by side, the experimental one included, on the three shapes
#   3. the bench line
TAG=${1:-r2a}
O=gpurun_out; mkdir -p $O
KCRN_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -rxXs > $O/${TAG}_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
{
for g in auto 0 1,2 3 1,3 3,2; do
  if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
  timeout 120 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -v "^$"
done
unset KCRN_GATHER
KCRN_GATHER_TRY_STAGED=1 KCRN_VERBOSE=1 timeout 120 python profiles/microbench/variant_time.py rcv1 w8a news20 2>&1 | grep -v "^$"
KCRN_GATHER_WIDE=1 KCRN_VERBOSE=1 timeout 120 python profiles/microbench/variant_time.py news20 2>&1 | grep -v "^$"
} > $O/${TAG}_variants.txt
cat $O/${TAG}_variants.txt
timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c1-400 $O/${TAG}_bench.json
