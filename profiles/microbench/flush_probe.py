"""The 256 MiB write that bench.py uses as its L2 flush leaves the recurrence ~9 % slower for good
(profiles/r3g_bench_vs_microbench.txt).  What exactly does, and what undoes it?"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench
torch.cuda.set_device(0)
from optimizer.loss import LogisticRegression

P = bench.make_shard("rcv1", 0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True, n_global=P.n)
eng = loss.engine
eng.set_x(loss.to_device(P.x0))
eng.gradient(keep_internal=True)


def rec(tag, m=10, reps=20, batches=3, warm=3):
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
    print(f"{tag:64s} {out} us per HVP", flush=True)


rec("fresh handle")
buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rec("256 MiB allocated, untouched")
s = int(buf[:: 4096].sum().item())
rec("256 MiB read sparsely (one byte per page)")
t = buf.view(torch.int64).sum().item()
rec("256 MiB read completely")
buf.fill_(1)
rec("256 MiB written (the bench's flush)")
rec("... 100 recurrences later", warm=100)
buf2 = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
t = buf2.view(torch.int64).sum().item()
rec("another 512 MiB read completely")
buf.fill_(2); torch.cuda.synchronize(); time.sleep(0.5)
rec("written again, 0.5 s pause")
buf.fill_(3); t = buf.view(torch.int64).sum().item()
rec("written, then read back")
del buf, buf2; torch.cuda.empty_cache()
rec("both buffers freed")
small = torch.empty(64 << 20, dtype=torch.uint8, device="cuda"); small.fill_(1)
rec("64 MiB written")
small.fill_(1); small.fill_(1)
rec("64 MiB written twice more")
