#!/bin/bash
TAG=${1:-r2x}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reorth or True or lanczos" > $O/${TAG}_reorth_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_reorth_pytest.log; tail -3 $O/${TAG}_reorth_pytest.log
timeout 300 python bench.py --no-strong --no-cpu-baseline > $O/${TAG}_bench_reorth.json 2>$O/${TAG}_bench_reorth.err
python -c "
import json; b=json.loads(open('$O/${TAG}_bench_reorth.json').read().strip().splitlines()[-1]); print(json.dumps(b['roofline']['reorth'], indent=0))"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-reps 5 --no-strong --reorth 1 --m 50"
timeout 300 $CMD > $O/${TAG}_reorth_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lanczos_update -s 300 -c 3 -o $O/${TAG}_ncu_reorth $CMD > $O/${TAG}_ncu_reorth.log 2>&1
echo "ncu exit $?"
