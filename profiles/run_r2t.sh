#!/bin/bash
# Round 2, session 2, first call: register-resident sub-problem kernel, stream words first, bulk-copy words (variant 3)
#   gpurun --timeout 900 -- 'bash profiles/run_r2t.sh r2t'
TAG=${1:-r2t}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_box.txt 2>&1
timeout 500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; RC=$?
echo "exit $RC" >> $O/${TAG}_pytest.log
tail -25 $O/${TAG}_pytest.log
{
for lib in default ilv; do
  if [ $lib = default ]; then unset KCRN_LIBRARY; else export KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_$lib.so; fi
  for g in auto 1,2 1,3 3; do
    if [ $g = auto ]; then unset KCRN_GATHER; else export KCRN_GATHER=$g; fi
    [ $lib = ilv ] && [ $g = 3 ] && continue
    echo "=== lib=$lib KCRN_GATHER=$g"
    KCRN_VERBOSE=1 timeout 100 python profiles/microbench/variant_time.py rcv1 w8a 2>&1 | grep -E "Lanczos|sweeps|pass, gather|Error|error"
  done
done
} > $O/${TAG}_variants.txt 2>&1
unset KCRN_LIBRARY KCRN_GATHER
grep -E "===|Lanczos graph" $O/${TAG}_variants.txt
timeout 250 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?" >> $O/${TAG}_bench.err
cut -c1-900 $O/${TAG}_bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-reps 5"
timeout 150 $CMD > $O/${TAG}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
grep -E "cubic" $O/${TAG}_ncu_launches.csv | head -5
