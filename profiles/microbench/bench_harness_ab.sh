#!/bin/bash
# Which part of the harness costs the host-driven step ~0.1 ms?  The same quick bench under every combination of: idle time
# behind the flush, lead time of the nvidia-smi sampler, step path (device line search / per-trial).
O=gpurun_out; TAG=${1:-r4c}
for ds in 1 0; do for lead in 0 1.5; do for sl in 0 2; do
  KCRN_BENCH_DEVICE_STEP=$ds KCRN_BENCH_SAMPLER_LEAD_S=$lead KCRN_BENCH_FLUSH_SLEEP_MS=$sl timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong --sweep-reps 10 2>/dev/null | python -c "
import json,sys
b=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('device_step=$ds sampler_lead=$lead flush_sleep_ms=$sl  ms/step %.4f  e2e %.4f  device_loop %.4f  in-graph us/HVP %.2f' % (b['ms_per_step'], b['e2e']['ms_per_step'], b['device_loop']['ms_per_step'], b['roofline']['kernel_us']))
"
done; done; done | tee $O/${TAG}_bench_harness_ab.txt
