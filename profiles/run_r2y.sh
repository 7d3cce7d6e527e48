#!/bin/bash
# K6 with the basis tile staged through shared memory: parity + timing per capacity
O=gpurun_out; mkdir -p $O; TAG=${1:-r2y}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged or reorth or lanczos or cubic" > $O/${TAG}_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
timeout 600 python profiles/microbench/reorth_time.py > $O/${TAG}_reorth_time.txt 2>&1
cat $O/${TAG}_reorth_time.txt
