#!/bin/bash
# K6 stages 1, 2 with the basis entries in registers: parity, C3-shape timing against the general kernel, per-launch times
O=gpurun_out; mkdir -p $O; TAG=${1:-r3d}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reorth or lanczos or trajectory" > $O/${TAG}_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
{
KCRN_NO_REORTH_REG=1 timeout 120 python profiles/microbench/c3_reorth.py child 2>&1 | tail -1 | sed 's/^/general kernel:  /'
timeout 120 python profiles/microbench/c3_reorth.py child 2>&1 | tail -1 | sed 's/^/register kernel: /'
KCRN_NO_REORTH_REG=1 timeout 120 python profiles/microbench/reorth_time.py 2>&1 | tail -3 | sed 's/^/general kernel:  /'
timeout 120 python profiles/microbench/reorth_time.py 2>&1 | tail -3 | sed 's/^/register kernel: /'
} > $O/${TAG}_reorth_time.txt 2>&1
cat $O/${TAG}_reorth_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lanczos_ -c 900 --csv --log-file $O/${TAG}_c3_reorth_launches.csv python profiles/microbench/c3_reorth.py child > $O/${TAG}_c3_ncu.log 2>&1
echo "ncu exit $?"
