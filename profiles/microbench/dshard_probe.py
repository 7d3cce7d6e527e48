"""d-sharded against replicated full re-orthogonalisation on a row-sharded problem (torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
       profiles/microbench/dshard_probe.py [n d nnz_per_row m]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import torch, torch.distributed as dist
from kcrn_b200.dist import join_communicator
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem_device

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, per_row, m = (int(float(a)) for a in (sys.argv[1:5] + ["19996", "1355191", "450", "50"][len(sys.argv) - 1:]))
n_loc = n // world
indptr, indices, data, b = make_problem_device(n_loc, d, float(per_row), seed=1000 + rank, perm_seed=7)
out = {"n_gpus": world, "n": n_loc * world, "d": d, "m": m}
x = torch.full((d,), 0.5, device="cuda", dtype=torch.float64)
keep = {}
for mode in ("0", "1"):
    os.environ["KCRN_DSHARD"] = mode
    eng = Engine(indptr, indices, data, b, (n_loc, d), n_global=n_loc * world)
    join_communicator(eng, rank, world)
    eng.set_x(x)
    eng.gradient(keep_internal=True)
    ms = {}
    for ro in (False, True):
        eng.lanczos(m, ro)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.lanczos(m, ro)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms[ro] = float(t.item())
    V, a, bt, info = eng.lanczos_view()
    keep[mode] = (a.cpu().numpy().copy(), V[:4].cpu().numpy().copy())
    out["dsharded" if mode == "1" else "replicated"] = {"recurrence_ms_reorth": ms[True], "recurrence_ms_plain": ms[False],
                                                        "reorth_ms": ms[True] - ms[False]}
    eng.close()
import numpy as np
out["alphas_rel_diff"] = float(np.linalg.norm(keep["1"][0] - keep["0"][0]) / np.linalg.norm(keep["0"][0]))
out["V_rel_diff"] = float(np.linalg.norm(keep["1"][1] - keep["0"][1]) / np.linalg.norm(keep["0"][1]))
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
