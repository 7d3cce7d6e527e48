"""Helper (not a test): one fused Lanczos recurrence on a named synthetic problem, results to an .npz.  Run as a
child process by the GPU tests so that process-wide switches of libkcrn (KCRN_NO_DEFER, KCRN_NO_PDL, KCRN_GATHER)
can be exercised:  python tests/lanczos_probe.py <problem> <m> <reorth> <out.npz>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import conftest  # noqa: F401  (import paths)
import torch
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem
from oracle import oracle as O

name, m, reorth, out = sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3])), sys.argv[4]
P = make_problem(name, seed=0)
eng = Engine(P.indptr, P.indices, P.data, O.map_labels(P.b), (P.n, P.d), l2=1e-4)
eng.set_x(torch.as_tensor(P.x0, device="cuda"))
rng = np.random.default_rng(5)
value, grad = eng.value(), eng.gradient().cpu().numpy()
hvp = eng.hvp(torch.as_tensor(rng.standard_normal(P.d), device="cuda")).cpu().numpy()
eng.gradient(keep_internal=True)
eng.lanczos(m, reorth)
V, al, bet, info = eng.lanczos_view()
np.savez(out, value=np.array([value]), grad=grad, hvp=hvp, V=V.cpu().numpy(), alphas=al.cpu().numpy(), bet=bet.cpu().numpy(), info=info.cpu().numpy(),
         variants=np.array(list(eng.gather_variants().values())))
eng.close()
