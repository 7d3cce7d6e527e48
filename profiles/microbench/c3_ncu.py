"""One Krylov-CRN trial at the C3 shape (news20: d = 1.36 M, m = 50) for an ncu launch list: which kernel owns the HVP."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench  # noqa: F401
torch.cuda.set_device(0)
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem_device
n, d, m = 19996, 1355191, 50
indptr, indices, data, b = make_problem_device(n, d, 450.0, seed=1000, perm_seed=7)
eng = Engine(indptr, indices, data, b, (n, d), n_global=n)
eng.set_x(torch.full((d,), 0.5, device="cuda", dtype=torch.float64))
r = eng.krylov_trial(True, m, False, 5e-4, 0.1, 1e-8, 100, x_out=eng.new_vector())
r = eng.krylov_trial(True, m, False, 2.5e-4, r[2], 1e-8, 100, x_out=eng.new_vector())
print("hvps", r[7])
