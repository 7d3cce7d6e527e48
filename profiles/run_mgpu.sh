#!/bin/bash
# multi-GPU checks at N ranks:  gpurun --gpus N --timeout 900 -- 'bash profiles/run_mgpu.sh <tag> N'
TAG=${1:-r2x}; N=${2:-2}
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_worker.py rcv1 > $O/${TAG}_mgpu_worker_N$N.log 2>&1
echo "worker exit $?"; grep -E "rank 0|OK|FAIL|Error|error" $O/${TAG}_mgpu_worker_N$N.log | tail -12 | cut -c1-300
bash profiles/run_bench.sh $TAG $N
