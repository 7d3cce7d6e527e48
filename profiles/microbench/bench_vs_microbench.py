"""Why does bench.py time the captured recurrence at ~22.4 us per HVP when variant_time.py sees ~20.4 on the same matrix?
The bench's sequence, with the recurrence timed after every stage."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
torch.cuda.set_device(0)
from optimizer.cubic import Cubic_Krylov_LS
from optimizer.loss import LogisticRegression

P = bench.make_shard("rcv1", 0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n)
eng = loss.engine


def rec(tag, m=10, reps=20, batches=3):
    eng.set_x(loss.to_device(P.x0))
    eng.gradient(keep_internal=True)
    for _ in range(3):
        eng.lanczos(m, False)
    out = []
    for _ in range(batches):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.lanczos(m, False)
        b.record(); torch.cuda.synchronize()
        out.append(round(1e3 * a.elapsed_time(b) / reps / m, 2))
    print(f"{tag:50s} {out} us per HVP, variants {tuple(eng.gather_variants().values())}", flush=True)


rec("fresh handle (LogisticRegression)")
opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=0, label="b", tqdm=False)
opt.init_run(P.x0)
for _ in range(8):
    eng.flush_l2(256 << 20)
    opt.step()
rec("after 8 optimizer steps, kcrn_flush_l2 before each")
v0 = loss.value(loss.to_device(P.x0))
eng.krylov_run(20, 10, False, v0, 1e-3, 0.5, 0.1)
loss.reset()
rec("after a device-loop run (20 iterations)")
opt2 = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=0, label="e", tqdm=False)
opt2.init_run(P.x0)
x_host = torch.as_tensor(P.x0).clone().pin_memory()
for _ in range(8):
    eng.flush_l2(256 << 20)
    opt2.x = x_host.to("cuda", non_blocking=True)
    opt2.step()
    x_host.copy_(opt2.x, non_blocking=True)
    torch.cuda.current_stream().synchronize()
rec("after 8 end-to-end steps (pinned x in / out)")
us = [eng.time_sweep(w, 200) for w in (0, 1)]
rec("after time_sweep")
print("torch allocator: allocated %.1f MB, reserved %.1f MB" % (torch.cuda.memory_allocated() / 1e6, torch.cuda.memory_reserved() / 1e6))
torch.cuda.empty_cache()
rec("after torch.cuda.empty_cache()")
print("torch allocator: allocated %.1f MB, reserved %.1f MB" % (torch.cuda.memory_allocated() / 1e6, torch.cuda.memory_reserved() / 1e6))
del opt, opt2
loss.reset()
torch.cuda.empty_cache()
rec("optimizers dropped, cache emptied")
print("torch allocator: allocated %.1f MB, reserved %.1f MB" % (torch.cuda.memory_allocated() / 1e6, torch.cuda.memory_reserved() / 1e6))
