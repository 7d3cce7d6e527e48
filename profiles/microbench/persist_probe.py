"""Persistent recurrence kernel (kcrn_persist.cu) against the graph of per-pass kernels on the same handle data:
   python profiles/microbench/persist_probe.py [shape ...]      (results + time per HVP, both paths)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"))
from kcrn_b200.engine import Engine          # noqa: E402
from kcrn_b200.synth import make_problem     # noqa: E402


def run(P, m, persist, l2=0.0, reps=20):
    os.environ["KCRN_PERSIST"] = "1" if persist else "0"
    b01 = (P.b + 1) / 2
    eng = Engine(P.indptr, P.indices, P.data, b01, (P.n, P.d), l2=l2)
    eng.set_x(torch.as_tensor(P.x0, device="cuda"))
    eng.gradient(keep_internal=True)
    for _ in range(3):
        eng.lanczos(m, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        for _ in range(reps):
            eng.lanczos(m, False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / reps)
    V, al, bet, info = eng.lanczos_view()
    out = dict(V=V.cpu().numpy().copy(), al=al.cpu().numpy().copy(), bet=bet.cpu().numpy().copy(),
               info=info.cpu().numpy().copy(), us=best)
    eng.close()
    return out


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


for name in (sys.argv[1:] or ["small", "w8a", "rcv1"]):
    P = make_problem(name, seed=0)
    for m, l2 in ((10, 0.0), (3, 1e-4)):
        t0 = time.time()
        a = run(P, m, True, l2)
        b = run(P, m, False, l2)
        hv = int(a["info"][3])
        k = int(min(a["info"][0], a["info"][1]))
        print(f"{name:6s} m={m:3d} l2={l2:g}: persistent {a['us'] / hv:7.2f} us/HVP   graph {b['us'] / max(1, int(b['info'][3])):7.2f} us/HVP   "
              f"info {a['info'].tolist()} vs {b['info'].tolist()}  alphas {rel(a['al'][:k], b['al'][:k]):.1e}  "
              f"bet {rel(a['bet'][:k], b['bet'][:k]):.1e}  V {rel(a['V'][:k], b['V'][:k]):.1e}   ({time.time() - t0:.0f}s)", flush=True)
