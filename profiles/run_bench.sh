#!/bin/bash
# bench.py as the driver runs it (N from $2), output to gpurun_out/<tag>_bench_N<N>.json
TAG=${1:-r2x}; N=${2:-1}; shift; shift
O=gpurun_out; mkdir -p $O
SECONDS=0
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > $O/${TAG}_bench_N1.json 2> $O/${TAG}_bench_N1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/${TAG}_bench_N$N.json 2> $O/${TAG}_bench_N$N.err
fi
echo "exit $? after ${SECONDS}s"; grep -vE "^\s" $O/${TAG}_bench_N$N.err | tail -5
python - <<PY
import json
try:
    b=json.loads([l for l in open("$O/${TAG}_bench_N$N.json") if l.startswith("{")][-1])
    print({k:b[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", b["e2e"]["value"], "devloop", (b.get("device_loop") or {}).get("ms_per_step"))
    print("roofline frac", b["roofline"]["frac"], b["roofline"]["kernel_us"], "reorth", json.dumps(b["roofline"].get("reorth")))
    print("parity", b.get("parity"))
    print("strong", json.dumps(b.get("strong"))[:1800])
    print("clocks", b["clocks"], "cpu", (b.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("parse failed", e)
PY
