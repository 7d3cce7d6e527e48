#!/bin/bash
# ncu --set full of selected kernels during the bench command:  bash profiles/run_ncu_full.sh <tag> <kernel-regex> <skip> <count> [bench args]
TAG=$1; RX=$2; SKIP=$3; CNT=$4; shift 4
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-reps 5 --no-strong $@"
timeout 200 $CMD > $O/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -o $O/${TAG}_ncu_full $CMD > $O/${TAG}_ncu2.log 2>&1
echo "ncu exit $?"; tail -3 $O/${TAG}_ncu2.log
