"""cProfile of SSCN steps (host-side overheads of the config-4 baseline):  python profiles/microbench/sscn_profile.py [m]"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "krylov-cubic-regularized-newton_b200"), ROOT]
import numpy as np, torch
from kcrn_b200.synth import make_problem
from optimizer.cubic import SSCN
from optimizer.loss import LogisticRegression
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10
P = make_problem("rcv1", seed=0)
loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
opt = SSCN(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, label="s", tqdm=False)
opt.rng = np.random.default_rng(42)
opt.init_run(P.x0)
for _ in range(5):
    opt.step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    opt.step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
