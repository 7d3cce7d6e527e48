"""BASELINE config 5 on one GPU: synthetic sparse logistic regression with power-law columns generated on the device
(default n = 10 M, d = 1 M, ~1e9 non-zeros), a few Krylov-CRN outer iterations, HVP time against the HBM roofline.
    python profiles/microbench/c5_run.py [n d nnz_per_row m steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import numpy as np, torch
from kcrn_b200.engine import Engine
from kcrn_b200.synth import make_problem_device

n, d, per_row, m, steps = (int(float(a)) for a in (sys.argv[1:6] + ["10e6", "1e6", "100", "100", "2"][len(sys.argv) - 1:]))
t0 = time.perf_counter()
indptr, indices, data, b = make_problem_device(n, d, float(per_row), seed=0)
torch.cuda.synchronize()
t_gen = time.perf_counter() - t0
nnz = int(indices.numel())
t0 = time.perf_counter()
eng = Engine(indptr, indices, data, b, (n, d))
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
del indptr, indices, data
torch.cuda.empty_cache()
x = torch.full((d,), 0.5, device="cuda", dtype=torch.float64)
eng.set_x(x)
us = [eng.time_sweep(w, 5) for w in (0, 1)]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = {"n": n, "d": d, "nnz": nnz, "m": m, "generate_s": t_gen, "create_s (upload + CSR->CSC + tiles)": t_build,
       "csr_pass_us": us[0], "csc_pass_us": us[1],
       "csr_frac": eng.sweep_bytes(0) / us[0] / 1e3 / peak, "csc_frac": eng.sweep_bytes(1) / us[1] / 1e3 / peak,
       "hvp_bytes": eng.hvp_bytes(), "hvp_frac_two_sweeps": eng.hvp_bytes() / (us[0] + us[1]) / 1e3 / peak}
# symmetry check of the HVP at full size
g = torch.Generator(device="cuda").manual_seed(1)
u, v = (torch.randn(d, device="cuda", dtype=torch.float64, generator=g) for _ in range(2))
Hu, Hv = eng.hvp(u), eng.hvp(v)
out["symmetry_rel"] = abs(float(u @ Hv) - float(v @ Hu)) / abs(float(u @ Hv))
# Krylov-CRN outer iterations
from optimizer.cubic import Cubic_Krylov_LS
class _Loss:                      # minimal oracle object around the existing engine (no second copy of the matrix)
    regularizer = None
    def __init__(s, e): s.engine, s.f_opt, s.x_opt, s.calls, s.n, s.dim = e, np.inf, None, dict(value=0, gradient=0, hess_vec_prod=0, forward=0), n, d
    def to_device(s, x): return x if hasattr(x, "data_ptr") else torch.as_tensor(x, device="cuda")
    def value(s, x): s.engine.set_x(x); v = s.engine.value(); s._note_value(v, x); return v
    def _forward(s, x): pass
    def adopt_iterate(s, x): pass
    def _note_value(s, v, x): s.f_opt = min(s.f_opt, v)
    def reset(s): pass
    @staticmethod
    def norm(x): return float(x.norm())
opt = Cubic_Krylov_LS(loss=_Loss(eng), reg_coef=1e-3, subspace_dim=m, tolerance=0, label="c5", tqdm=False)
opt.init_run(x)
opt.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    opt.step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
hv = sum(h["hvps"] for h in opt.history[1:])
out.update({"outer_iters_per_sec": steps / dt, "hvps_per_sec": hv / dt, "hvp_us_in_step": 1e6 * dt / hv,
            "hvp_frac_in_step": eng.hvp_bytes() / (1e6 * dt / hv) / 1e3 / peak, "f": [h["f"] for h in opt.history],
            "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9})
print(json.dumps(out))
