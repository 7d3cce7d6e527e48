"""kcrn_cubic_dense (one kernel) against the cuSOLVER-driven Newton loop on SSCN-like blocks: iterations, r, s, time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import numpy as np, torch
import optimizer.dense as D
from kcrn_b200.synth import make_problem
from optimizer.loss import LogisticRegression
P = make_problem("rcv1", seed=0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
x = torch.as_tensor(P.x0, device="cuda")
rng = np.random.default_rng(42)
eps = np.finfo(float).eps
for m in (10, 50):
    I = rng.choice(P.d, size=m, replace=False)
    g = loss.partial_gradient(x, I)
    H = loss.partial_hessian(x, I)
    for M in (5e-4, 1e-6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        a = D.cubic_solver_dense(g, H, M, r0=0.1, epsilon=eps)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        keep = D.DENSE_KERNEL_MAX; D.DENSE_KERNEL_MAX = 0
        b = D.cubic_solver_dense(g, H, M, r0=0.1, epsilon=eps)
        D.DENSE_KERNEL_MAX = keep
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"m={m} M={M:g}: kernel its {a[1]} r {a[2]:.17g} dec {a[3]:.6e} ({1e3 * (t1 - t0):.2f} ms) | torch its {b[1]} r {b[2]:.17g} dec {b[3]:.6e} "
              f"({1e3 * (t2 - t1):.2f} ms) | s rel {float((a[0] - b[0]).norm() / b[0].norm()):.2e}")
