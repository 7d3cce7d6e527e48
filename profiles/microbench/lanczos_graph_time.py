"""Time of the captured m-step Lanczos graph per HVP (rcv1 shape), for A/B experiments on the kernels."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import conftest  # noqa
import numpy as np, torch
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem
from oracle import oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "rcv1"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
P = make_problem(name, seed=0)
eng = Engine(P.indptr, P.indices, P.data, O.map_labels(P.b), (P.n, P.d))
eng.set_x(torch.as_tensor(P.x0, device="cuda"))
eng.gradient(keep_internal=True)
for _ in range(5):
    eng.lanczos(m, False)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    a.record()
    for _ in range(50):
        eng.lanczos(m, False)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 50 * 1e3)
hv = int(eng.lanczos_view()[3][3].item()) or m
print(f"{name} m={m}: graph {best:.1f} us, {hv} HVPs, {best / hv:.2f} us per HVP  env={ {k: v for k, v in os.environ.items() if k.startswith('KCRN_')} }")
