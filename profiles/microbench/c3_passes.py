"""The two passes of an HVP at the C3 shape (news20: n = 19 996, d = 1 355 191, 7.0 M non-zeros), each on its own and
inside the recurrence: which one owns the 117 us per HVP, per forced gather variant."""
import os, sys, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench  # noqa: F401  (puts the package directory on sys.path)
torch.cuda.set_device(0)
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem_device

n, d, m = 19996, 1355191, 50
indptr, indices, data, b = make_problem_device(n, d, 450.0, seed=1000, perm_seed=7)
nnz = int(indices.numel())
eng = Engine(indptr, indices, data, b, (n, d), n_global=n)
x = torch.full((d,), 0.5, device="cuda", dtype=torch.float64)
eng.set_x(x)
eng.gradient(keep_internal=True)
us = [eng.time_sweep(w, 50) for w in (0, 1)]
by = [eng.sweep_bytes(w) for w in (0, 1)]
alt = eng.time_sweep(2, 50)
print(f"plain passes alternating (one plain HVP = CSR pass + CSC pass, the matrices take turns in L2): {2 * alt:.1f} us per HVP", flush=True)
print(f"nnz {nnz}  plain passes: CSR {us[0]:.1f} us ({by[0] / us[0] / 1e3:.0f} GB/s)  CSC {us[1]:.1f} us ({by[1] / us[1] / 1e3:.0f} GB/s)  variants {eng.gather_variants()}", flush=True)
# a recurrence from a point where it does not break down early: one Newton-ish step away from x0
from optimizer.cubic import Cubic_Krylov_LS  # noqa: F401  (same import path as the bench)
val, dec, r, its, status, m_eff, gnorm, hv, dx = eng.krylov_trial(True, m, False, 5e-4, 0.1, 1e-8, 100, x_out=eng.new_vector())
eng.gradient(keep_internal=True)
for _ in range(2):
    eng.lanczos(m, False)
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(4):
    eng.lanczos(m, False)
c.record(); torch.cuda.synchronize()
hv = int(eng.lanczos_view()[3][3].item())
print(f"recurrence m = {m}: {hv} HVPs, {1e3 * a.elapsed_time(c) / 4 / max(1, hv):.1f} us per HVP; HVP bytes {eng.hvp_bytes()}", flush=True)
