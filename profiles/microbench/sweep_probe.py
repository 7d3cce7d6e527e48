"""Probe for ncu: builds the rcv1-shape engine and runs a few launches of each sweep kernel and one Lanczos graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import numpy as np, torch
from kcrn_b200.synth import make_problem
from kcrn_b200.engine import Engine
from oracle.oracle import map_labels

name = sys.argv[1] if len(sys.argv) > 1 else "rcv1"
P = make_problem(name, seed=0)
eng = Engine(P.indptr, P.indices, P.data, map_labels(P.b), (P.n, P.d))
x = torch.as_tensor(P.x0, device="cuda")
eng.set_x(x)
eng.gradient(keep_internal=True)
for which in (0, 1):
    print("sweep", which, "us", eng.time_sweep(which, 3))
eng.lanczos(10, False)
torch.cuda.synchronize()
print(eng.counters())
# fused vs step-wise agreement
from optimizer.cubic import Lanczos
g = eng.gradient()
eng.lanczos(8, False, g=g)
V, al, bet, info = eng.lanczos_view()
V2, al2, be2, _ = Lanczos(lambda v: eng.hvp(v), g, m=8)
print("alphas fused   ", al.cpu().numpy())
print("alphas stepwise", al2.cpu().numpy())
print("betas fused    ", bet.cpu().numpy())
print("betas stepwise ", be2.cpu().numpy(), "info", info.cpu().numpy())
