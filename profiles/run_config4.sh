#!/bin/bash
TAG=${1:-r2x}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_config4.py -m gpu -q > $O/${TAG}_config4_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_config4_pytest.log; tail -5 $O/${TAG}_config4_pytest.log
timeout 600 python bench_config4.py --crn-iters 1 > $O/${TAG}_bench_config4.json 2> $O/${TAG}_bench_config4.err
echo "bench exit $?"; tail -3 $O/${TAG}_bench_config4.err; cat $O/${TAG}_bench_config4.json
# FP64 tensor-pipe utilisation of the dense A'DA kernel (config 4; north_star asks for the ncu capture)
CMD="python bench_config4.py --skip-crn --sscn-iters 1"
timeout 300 $CMD > $O/${TAG}_c4_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_tn -s 1 -c 1 -o $O/${TAG}_ncu_dgemm $CMD > $O/${TAG}_ncu_dgemm.log 2>&1
echo "ncu exit $?"
