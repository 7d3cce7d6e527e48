"""Which flush procedure leaves the recurrence in its fast state (20.6 us per HVP at the rcv1 shape) and which in the slow
one (21.9 - 22.5)?  Each procedure is followed by the same thing: optionally one optimizer step, then the timed recurrences."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench
torch.cuda.set_device(0)
from optimizer.cubic import Cubic_Krylov_LS
from optimizer.loss import LogisticRegression

P = bench.make_shard("rcv1", 0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n)
eng = loss.engine
opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=0, label="b", tqdm=False)
opt.init_run(P.x0)
for _ in range(4):
    opt.step()


def rec(tag, m=10, reps=20, batches=2, warm=3):
    eng.set_x(loss.to_device(P.x0))
    eng.gradient(keep_internal=True)
    for _ in range(warm):
        eng.lanczos(m, False)
    out = []
    for _ in range(batches):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.lanczos(m, False)
        b.record(); torch.cuda.synchronize()
        out.append(round(1e3 * a.elapsed_time(b) / reps / m, 2))
    print(f"{tag:72s} {out} us per HVP", flush=True)


def torch_flush(write=True):
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if write:
        buf.fill_(1)
    else:
        buf.view(torch.int64).sum().item()
    torch.cuda.synchronize()
    del buf
    torch.cuda.empty_cache()


rec("control: 4 optimizer steps, no flush")
opt.step(); rec("control: one more step")
eng.flush_l2(256 << 20); rec("kcrn_flush_l2 (write 256 MiB, read 256 MiB, free both), recurrences right away")
eng.flush_l2(256 << 20); opt.step(); rec("kcrn_flush_l2, one optimizer step")
torch_flush(True); rec("torch buffer written, freed (empty_cache), recurrences right away")
torch_flush(True); opt.step(); rec("torch buffer written, freed, one optimizer step")
torch_flush(False); opt.step(); rec("torch buffer READ, freed, one optimizer step")
eng.flush_l2(256 << 20); eng.flush_l2(256 << 20); opt.step(); rec("kcrn_flush_l2 twice, one optimizer step")
eng.flush_l2(1 << 30); opt.step(); rec("kcrn_flush_l2 of 1 GiB, one optimizer step")
torch_flush(True); rec("torch buffer written, freed, recurrences right away (again)")
for _ in range(3):
    opt.step()
rec("three more optimizer steps, no flush")
