#!/bin/bash
# One gpurun call that refreshes the judged evidence at the current HEAD:
#   gpurun --timeout 840 -- 'bash profiles/run_gpu_round.sh r1d'
# writes gpurun_out/<tag>_{pytest.log,bench.json,bench_reference_arm.json,ncu_launches.csv,ncu_full.ncu-rep}
TAG=${1:-r1x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_box.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; RC=$?
echo "exit $RC" >> $O/${TAG}_pytest.log
if [ $RC -ne 0 ]; then tail -40 $O/${TAG}_pytest.log; exit 1; fi      # nothing below is worth box time on a red suite
timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?" >> $O/${TAG}_bench.err
timeout 120 python bench.py --impl reference --steps 10 --warmup 2 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err
# ncu: the same bench command, short (the plain run right before it must exit 0)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-reps 5"
timeout 150 $CMD > $O/${TAG}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
timeout 150 $CMD > $O/${TAG}_plain2.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 400 -c 6 -o $O/${TAG}_ncu_full $CMD > $O/${TAG}_ncu2.log 2>&1
tail -3 $O/${TAG}_pytest.log
cat $O/${TAG}_bench.json | cut -c1-600
# measured choice of the gather variant against the forced ones (same box)
{
for g in auto 0 1,2; do
  if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
  timeout 100 python profiles/microbench/variant_time.py ${SHAPES:-rcv1 w8a} 2>&1 | grep -v "^$"
done
} > $O/${TAG}_variants.txt
cat $O/${TAG}_variants.txt
