#!/bin/bash
# persistent recurrence kernel: probe with the per-phase trace, then (optionally) tests + bench
TAG=${1:-r2d}
O=gpurun_out; mkdir -p $O
KCRN_PERSIST_TRACE=1 timeout 300 python profiles/microbench/persist_probe.py rcv1 > $O/${TAG}_probe.txt 2>&1
echo "exit $?" >> $O/${TAG}_probe.txt
grep -E "persistent  |trace|mean|one row" $O/${TAG}_probe.txt | cut -c1-200
if [ "$2" = "full" ] && grep -q "exit 0" $O/${TAG}_probe.txt; then
  timeout 600 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1
  echo "exit $?" >> $O/${TAG}_pytest.log
  grep -E "TRAJ|passed|failed|FAILED|news20 state" $O/${TAG}_pytest.log | cut -c1-300
  timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
  cut -c1-600 $O/${TAG}_bench.json
fi
