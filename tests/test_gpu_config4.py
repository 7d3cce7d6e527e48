"""GPU parity for the config-4 rows (full-Hessian CRN, SSCN): dense A' D A on the FP64 tensor cores, the
column-block gradient / Hessian, the incremental A x update, the dense cubic solve and the two baseline
optimizers -- against golden fixtures written from the unmodified reference and against the CPU oracle."""
import functools

import numpy as np
import pytest

from conftest import rel_err
from kcrn_b200.synth import make_problem as _make_problem
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@functools.lru_cache(maxsize=None)
def make_problem(name, seed=0):
    return _make_problem(name, seed=seed)


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.fail("gpu-marked test on a machine without CUDA")
    return t


def _loss(P, l2=0.0):
    from optimizer.loss import LogisticRegression
    return LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=l2, store_mat_vec_prod=True)


@pytest.mark.parametrize("tag", ["tiny", "tiny_l2"])
def test_dense_hessian_and_blocks_vs_reference_golden(torch, golden, tag):
    g = golden("oracle_methods")
    P = make_problem("tiny", seed=3)
    loss = _loss(P, float(g[f"{tag}_l2"]))
    x, I = g[f"{tag}_x"], g[f"{tag}_I"]
    assert rel_err(loss.hessian(x), g[f"{tag}_hess"]) <= 1e-12
    assert rel_err(loss.partial_gradient(x, I), g[f"{tag}_pgrad"]) <= 1e-12
    assert rel_err(loss.partial_hessian(x, I), g[f"{tag}_phess"]) <= 1e-12
    Ax = loss.mat_vec_product(x).copy()
    loss.update_mat_vec_product(Ax, g[f"{tag}_delta"], I)
    assert rel_err(loss._mat_vec_prod.cpu().numpy(), g[f"{tag}_Ax_upd"]) <= 1e-12
    # the override holds until reset(): value() now uses the updated A x, whatever x is (loss.py:267, 279-286)
    L = O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b, l2=float(g[f"{tag}_l2"]))
    Axu = g[f"{tag}_Ax_upd"]
    want = np.mean((1 - L.b) * Axu - O.logsig(Axu)) + float(g[f"{tag}_l2"]) / 2 * np.linalg.norm(x) ** 2
    assert abs(loss.value(x) - want) <= 1e-12 * abs(want)
    loss.reset()
    assert abs(loss.value(x) - float(g[f"{tag}_value"])) <= 1e-12 * abs(float(g[f"{tag}_value"]))


def test_dense_hessian_matches_hvp_and_oracle_mid_size(torch):
    # d = 500 needs 4 x 4 output tiles (one ragged), n = 3000 -> 188 k-chunks: exercises the DMMA pipeline
    P = make_problem("small", seed=0)
    loss = _loss(P, 1e-3)
    x = torch.as_tensor(P.x0 + 0.1 * np.random.default_rng(0).standard_normal(P.d), device="cuda")
    H = loss.hessian(x)
    assert H.shape == (P.d, P.d)
    assert float((H - H.T).abs().max()) <= 1e-15
    v = torch.as_tensor(np.random.default_rng(1).standard_normal(P.d), device="cuda")
    hv = loss.hess_vec_prod(x, v)
    assert rel_err((H @ v).cpu().numpy(), hv.cpu().numpy()) <= 1e-12
    L = O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b, l2=1e-3)
    assert rel_err(H.cpu().numpy(), L.hessian(x.cpu().numpy())) <= 1e-12
    I = np.random.default_rng(42).choice(P.d, size=130, replace=False)      # 130 -> two column tiles
    assert rel_err(loss.partial_hessian(x, I).cpu().numpy(), L.partial_hessian(x.cpu().numpy(), I)) <= 1e-12
    assert rel_err(loss.partial_gradient(x, I).cpu().numpy(), L.partial_gradient(x.cpu().numpy(), I)) <= 1e-12


def test_dense_cubic_solver_vs_reference_golden(torch, golden):
    from optimizer.cubic import cubic_solver_root
    g = golden("cubic")
    Md, r0d, r_ref, dec_ref = g["dense_scal"]
    s, its, r, dec = cubic_solver_root(g["dense_g"], g["dense_H"], Md, r0=r0d)
    assert its == int(g["dense_its"]) and rel_err(s, g["dense_s"]) <= 1e-11
    assert abs(r - r_ref) <= 1e-12 * r_ref and abs(dec - dec_ref) <= 1e-11 * abs(dec_ref)


@pytest.mark.parametrize("m", [10, 50])
def test_sscn_trajectory_vs_reference(torch, golden, m):
    from optimizer.cubic import SSCN
    g = golden("traj_config4")
    P = make_problem("small", seed=0)
    assert P.fingerprint() == str(g["fp"])
    opt = SSCN(loss=_loss(P), reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, label="s", tqdm=False)
    ref = g[f"sscn{m}_f"]
    f = []
    opt.rng = np.random.default_rng(42)
    opt.init_run(P.x0)
    for _ in range(len(ref)):
        opt.step()
        f.append(opt.value)
    f = np.array(f)
    env = np.abs(g[f"sscn{m}_rev_f"] - ref) / np.abs(ref)
    dev = np.abs(f - ref) / np.abs(ref)
    print(f"\nSSCN m={m}: max dev {dev.max():.2e} (reference self-noise {env.max():.2e})")
    assert dev.max() <= 1e-8
    assert opt.reg_coef == g[f"sscn{m}_M"][-1]
    # SSCN asks Newton for |p - p0| <= machine epsilon (cubic.py:374): the last iterations move p by an ulp or
    # two, so the count depends on the rounding of the Cholesky solves (cuSOLVER here, LAPACK there)
    want_its = int(g[f"sscn{m}_solver_it"][-1])
    assert abs(opt.solver_it - want_its) <= 0.05 * want_its
    assert rel_err(opt.x.cpu().numpy(), g[f"sscn{m}_x_final"]) <= 1e-8


def test_sscn_trajectory_with_l2_vs_reference(torch, golden):
    """l2 > 0: value and block gradient depend on x itself; the cached-A x override (update_mat_vec_product, loss.py:279-286)
    must not freeze the x the engine uses for those terms (kcrn_set_x_only)."""
    from optimizer.cubic import SSCN
    from optimizer.loss import LogisticRegression
    g = golden("traj_config4_l2")
    P = make_problem("small", seed=0)
    assert P.fingerprint() == str(g["fp"])
    loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=float(g["l2"]), store_mat_vec_prod=True)
    opt = SSCN(loss=loss, reg_coef=1e-3, subspace_dim=10, tolerance=1e-9, label="s", tqdm=False)
    opt.rng = np.random.default_rng(42)
    opt.init_run(P.x0)
    ref, f = g["sscn10_f"], []
    for _ in range(len(ref)):
        opt.step()
        f.append(opt.value)
    dev = np.abs(np.array(f) - ref) / np.abs(ref)
    print(f"\nSSCN m=10, l2=1e-3: max dev {dev.max():.2e}")
    assert dev.max() <= 1e-8
    assert opt.reg_coef == g["sscn10_M"][-1]
    assert rel_err(opt.x.cpu().numpy(), g["sscn10_x_final"]) <= 1e-8


def test_full_hessian_crn_trajectory_vs_reference(torch, golden):
    from optimizer.cubic import Cubic_LS
    g = golden("traj_config4")
    P = make_problem("small", seed=0)
    opt = Cubic_LS(loss=_loss(P), reg_coef=1e-3, cubic_solver="full", tolerance=1e-8, label="c", tqdm=False)
    opt.init_run(P.x0)
    ref = g["full_f"]
    f = []
    for _ in range(len(ref)):
        opt.step()
        f.append(opt.value)
    dev = np.abs(np.array(f) - ref) / np.abs(ref)
    print(f"\nfull-Hessian CRN: max dev {dev.max():.2e}")
    assert dev.max() <= 1e-8
    assert opt.reg_coef == g["full_M"][-1] and opt.solver_it == int(g["full_solver_it"][-1])
    assert rel_err(opt.x.cpu().numpy(), g["full_x_final"]) <= 1e-8


def test_cg_crn_trajectory_vs_reference(torch, golden):
    """Cubic_LS(cubic_solver="CG") -- the CRN baseline the reference driver actually runs for dim >= 500."""
    from optimizer.cubic import Cubic_LS
    g = golden("traj_config4")
    P = make_problem("small", seed=0)
    opt = Cubic_LS(loss=_loss(P), reg_coef=1e-3, cubic_solver="CG", tolerance=1e-8, label="c", tqdm=False)
    opt.init_run(P.x0)
    ref = g["cg_f"]
    f = []
    for _ in range(len(ref)):
        opt.step()
        f.append(opt.value)
    dev = np.abs(np.array(f) - ref) / np.abs(ref)
    print(f"\nCG CRN: max dev {dev.max():.2e}, {opt.cg_iterations} CG iterations")
    assert dev.max() <= 1e-8
    assert opt.reg_coef == g["cg_M"][-1] and opt.solver_it == int(g["cg_solver_it"][-1])
    assert rel_err(opt.x.cpu().numpy(), g["cg_x_final"]) <= 1e-7
