#!/bin/bash
# GPU suite + bench at HEAD:  gpurun --timeout 900 -- 'bash profiles/run_gpu.sh <tag> [pytest-args]'
TAG=${1:-r2x}; shift
O=gpurun_out; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -q -s "$@" > $O/${TAG}_pytest.log 2>&1
echo "exit $?" >> $O/${TAG}_pytest.log
grep -E "^TRAJ|news20 state|passed|failed|FAILED|Error|exit" $O/${TAG}_pytest.log | cut -c1-330
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; tail -3 $O/${TAG}_bench.err
python - <<PY
import json
try:
    b=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print({k:b[k] for k in ("value","ms_per_step","hvps_per_step","gpu_launches")}, b["e2e"], b.get("device_loop"))
    print(json.dumps(b["roofline"])[:900])
    print(b.get("cpu_baseline"))
except Exception as e: print("bench parse failed", e)
PY
