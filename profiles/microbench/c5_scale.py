"""BASELINE config 5, row-sharded over the GPUs of one box (strong scaling of the 1e9-non-zero problem):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        profiles/microbench/c5_scale.py [n d nnz_per_row m]
Every rank generates its own block of n / N rows on the device; HVP = local sweeps + one NCCL allreduce of the
d-vector.  Times are CUDA events, max over ranks."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import torch, torch.distributed as dist
from kcrn_b200.dist import join_communicator
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem_device

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, per_row, m = (int(float(a)) for a in (sys.argv[1:5] + ["10e6", "1e6", "100", "100"][len(sys.argv) - 1:]))
n_loc = n // world
indptr, indices, data, b = make_problem_device(n_loc, d, float(per_row), seed=rank)
nnz = int(indices.numel())
eng = Engine(indptr, indices, data, b, (n_loc, d), n_global=n_loc * world)
del indptr, indices, data
torch.cuda.empty_cache()
if world > 1:
    join_communicator(eng, rank, world)
x = torch.full((d,), 0.5, device="cuda", dtype=torch.float64)
v = torch.randn(d, device="cuda", dtype=torch.float64, generator=torch.Generator(device="cuda").manual_seed(7))
eng.set_x(x)
out = torch.empty_like(v)
for _ in range(3):
    eng.hvp(v, out)

def timed(fn, reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

hvp_ms = timed(lambda: eng.hvp(v, out), 10)
eng.gradient(keep_internal=True)
eng.lanczos(m, False)
lan_ms = timed(lambda: eng.lanczos(m, False), 2)
tot = torch.tensor([nnz], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot)
if rank == 0:
    print(json.dumps({"n_gpus": world, "n": n_loc * world, "d": d, "nnz_total": int(tot.item()), "nnz_rank0": nnz,
                      "hvp_ms": hvp_ms, "hvps_per_sec": 1e3 / hvp_ms, "lanczos_m": m, "lanczos_ms": lan_ms,
                      "hvp_ms_inside_lanczos": lan_ms / m, "hvps_per_sec_inside_lanczos": 1e3 * m / lan_ms}), flush=True)
eng.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
