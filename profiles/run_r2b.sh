#!/bin/bash
# persistent recurrence kernel: first run.  gpurun --timeout 600 -- 'bash profiles/run_r2b.sh r2b'
TAG=${1:-r2b}
O=gpurun_out; mkdir -p $O
KCRN_VERBOSE=1 timeout 300 python profiles/microbench/persist_probe.py small > $O/${TAG}_probe.txt 2>&1
echo "exit $?" >> $O/${TAG}_probe.txt
tail -12 $O/${TAG}_probe.txt
if grep -q "exit 0" $O/${TAG}_probe.txt; then
  timeout 300 python profiles/microbench/persist_probe.py w8a rcv1 >> $O/${TAG}_probe.txt 2>&1
  echo "exit $?" >> $O/${TAG}_probe.txt
  tail -6 $O/${TAG}_probe.txt
  timeout 400 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
  echo "exit $?" >> $O/${TAG}_pytest.log
  tail -15 $O/${TAG}_pytest.log
  timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
  cut -c1-600 $O/${TAG}_bench.json
fi
