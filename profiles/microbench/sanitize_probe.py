"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): create, all sweeps, fused Lanczos (both
modes), sub-problem, trial point, step record, config-4 kernels -- on the 'tiny' and 'small' problems."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import numpy as np, torch
from kcrn_b200.synth import make_problem
from optimizer.cubic import Cubic_Krylov_LS, SSCN
from optimizer.loss import LogisticRegression

for name, m in (("tiny", 4), ("small", 6)):
    P = make_problem(name, seed=3)
    loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=1e-3, store_mat_vec_prod=True)
    for reorth in (False, True):
        opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, reorth=reorth, label="s", tqdm=False)
        opt.run(P.x0, it_max=3)
    x = torch.as_tensor(P.x0, device="cuda")
    H = loss.hessian(x)
    I = np.arange(0, P.d, 7)[:8]
    loss.partial_gradient(x, I); loss.partial_hessian(x, I)
    s = SSCN(loss=loss, reg_coef=1e-3, subspace_dim=5, label="q", tqdm=False)
    s.run(P.x0, it_max=2)
    print(name, "ok", float(H.trace()), opt.value)
torch.cuda.synchronize()
