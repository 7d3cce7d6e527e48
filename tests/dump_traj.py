"""Helper (not a test): dump engine trajectories for every golden case into gpurun_out/traj_dump.npz."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(__file__))
import conftest  # noqa
from kcrn_b200.synth import make_problem
from optimizer.cubic import Cubic_Krylov_LS
from optimizer.loss import LogisticRegression

out = {}
for name, ms, iters in (("small", (10,), 30), ("w8a", (10,), 30), ("rcv1", (10, 50), 40)):
    P = make_problem(name, seed=0)
    for m in ms:
        for reorth in (False, True):
            loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
            opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, reorth=reorth, label="k", tqdm=False)
            opt.run(P.x0, it_max=iters)
            for key in ("f", "M", "m_eff", "solver_it", "backtracks", "dx"):
                out[f"{name}_m{m}_{'ro' if reorth else 'ref'}_{key}"] = np.array([h[key] for h in opt.history], dtype=np.float64)
np.savez(os.path.join(conftest.ROOT, "gpurun_out", "traj_dump.npz"), **out)
print("dumped", len(out))
