#!/bin/bash
# Round-end evidence at HEAD, one call:  gpurun --timeout 1700 -- 'bash profiles/run_final.sh r3f'
TAG=${1:-r3f}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_box.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; RC=$?
echo "exit $RC" >> $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|exit" $O/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err; echo "reference arm exit $?"
timeout 600 python bench_config4.py --crn-iters 1 > $O/${TAG}_bench_config4.json 2> $O/${TAG}_bench_config4.err; echo "config4 exit $?"
python - <<PY
import json
b=json.loads([l for l in open("$O/${TAG}_bench.json") if l.startswith("{")][-1])
print({k:b[k] for k in ("value","ms_per_step","hvps_per_step","gpu_launches")}, "e2e", round(b["e2e"]["value"]), "devloop", b["device_loop"].get("ms_per_step"))
print("roofline", round(b["roofline"]["frac"],4), round(b["roofline"]["kernel_us"],2), "plain", b["roofline"]["other"]["kernel_us"], b["roofline"]["other"]["frac"])
print("reorth", json.dumps(b["roofline"]["reorth"])[:600])
print("strong", json.dumps(b["strong"])[:1800])
print("cpu", b["cpu_baseline"]["value"], "clocks", b["clocks"])
r=json.loads([l for l in open("$O/${TAG}_bench_reference_arm.json") if l.startswith("{")][-1]); print("reference arm", r["value"], r["ms_per_step"])
c=json.loads([l for l in open("$O/${TAG}_bench_config4.json") if l.startswith("{")][-1]); print("config4", c["dense_hessian"]["tflops_issued"], c["sscn_iters_per_sec"], c.get("full_crn"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-reps 5 --no-strong"
timeout 150 $CMD > $O/${TAG}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
echo "ncu launch list exit $?"
if [ -z "$SKIP_SWEEP_NCU" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 400 -c 6 -o $O/${TAG}_ncu_full_sweeps $CMD > $O/${TAG}_ncu2.log 2>&1
echo "ncu full (sweeps) exit $?"
fi
timeout 120 python profiles/microbench/c3_reorth.py child > $O/${TAG}_c3_reorth.txt 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lanczos_reorth_reg -s 150 -c 3 -o $O/${TAG}_ncu_full_reorth python profiles/microbench/c3_reorth.py child > $O/${TAG}_ncu3.log 2>&1
echo "ncu full (reorth) exit $?"; tail -1 $O/${TAG}_c3_reorth.txt
ls -la $O | grep ${TAG}
