"""GPU, larger than anything the CPU oracle handles in seconds: size-independent properties of the HVP and an
independent check against cuSPARSE (torch.sparse, test-only) on a ~50 M non-zero power-law problem generated on
the device (BASELINE config 5 recipe at 1/20 scale; the 1e9 run is profiles/microbench/c5_run.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a machine without CUDA")
    from kcrn_b200.engine import Engine
    from kcrn_b200.synth import make_problem_device
    n, d = 1_000_000, 200_000
    indptr, indices, data, b = make_problem_device(n, d, 50.0, seed=1)
    eng = Engine(indptr, indices, data, b, (n, d), l2=1e-5)
    A = torch.sparse_csr_tensor(indptr.to(torch.int64), indices.to(torch.int64), data, size=(n, d))
    return torch, eng, A, b, (n, d)


def test_csc_copy_matches_cusparse_transpose(big):
    torch, eng, A, b, (n, d) = big
    cp, ci, cx = eng.csc()
    C = A.to_sparse_csc()
    assert np.array_equal(cp, C.ccol_indices().cpu().numpy().astype(np.int32))
    assert np.array_equal(ci, C.row_indices().cpu().numpy().astype(np.int32))
    assert np.array_equal(cx, C.values().cpu().numpy())


def test_value_gradient_hvp_vs_cusparse(big):
    torch, eng, A, b, (n, d) = big
    g = torch.Generator(device="cuda").manual_seed(3)
    x = 0.3 * torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
    v = torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
    eng.set_x(x)
    Ax = A @ x
    a = torch.sigmoid(Ax)
    f = torch.mean((1 - b) * Ax - torch.nn.functional.logsigmoid(Ax)) + 1e-5 / 2 * (x @ x)
    grad = (A.t() @ (a - b)) / n + 1e-5 * x
    hv = (A.t() @ (a * (1 - a) * (A @ v))) / n + 1e-5 * v
    rel = lambda p, q: float((p - q).norm() / q.norm())
    assert abs(eng.value() - float(f)) <= 1e-12 * abs(float(f))
    assert rel(eng.cached_Ax(), Ax) <= 1e-12
    assert rel(eng.gradient(), grad) <= 1e-12
    assert rel(eng.hvp(v), hv) <= 1e-12


def test_hvp_symmetry_linearity_determinism(big):
    torch, eng, A, b, (n, d) = big
    g = torch.Generator(device="cuda").manual_seed(4)
    x = 0.3 * torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
    u = torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
    v = torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
    eng.set_x(x)
    Hu, Hv = eng.hvp(u), eng.hvp(v)
    assert torch.equal(Hu, eng.hvp(u))
    s1, s2 = float(u @ Hv), float(v @ Hu)
    assert abs(s1 - s2) <= 1e-12 * max(abs(s1), abs(s2))
    lin = eng.hvp(0.5 * u - 2.0 * v)
    assert float((lin - (0.5 * Hu - 2.0 * Hv)).norm() / lin.norm()) <= 1e-13
    assert float(u @ Hu) > 0                       # positive definite (l2 > 0)


def test_lanczos_reorth_orthonormal_basis_and_ritz_bounds(big):
    torch, eng, A, b, (n, d) = big
    x = torch.zeros(d, device="cuda", dtype=torch.float64)
    eng.set_x(x)
    eng.gradient(keep_internal=True)
    m = 12
    eng.lanczos(m, reorth=True)
    V, al, bet, info = eng.lanczos_view()
    assert int(info[0]) == m
    G = (V @ V.T).cpu().numpy()
    assert np.max(np.abs(G - np.eye(m))) <= 1e-12
    # T = V' H V: the tridiagonal the recurrence reports agrees with the projected operator
    HV = torch.stack([eng.hvp(V[i].contiguous()) for i in range(m)])
    T = (V @ HV.T).cpu().numpy()
    Tt = np.diag(al.cpu().numpy()) + np.diag(bet[1:m].cpu().numpy(), 1) + np.diag(bet[1:m].cpu().numpy(), -1)
    assert np.max(np.abs(T - Tt)) <= 1e-10 * np.max(np.abs(Tt))


def test_c5_recipe_at_one_percent_vs_oracle():
    """BASELINE configs[4] (n = 10 M, d = 1 M, ~116 non-zeros per row, generated on the device) at 1/100 of its rows and
    columns -- the largest instance of that recipe the CPU oracle finishes in seconds: value, A x, gradient and HVP of the
    engine against the oracle's scipy-order arithmetic on the SAME matrix (copied to the host), 1e-12 as for every other
    shape; the CSC copy bit for bit against the oracle's csr_tocsc; two Krylov-CRN iterations with m = 100."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a machine without CUDA")
    from kcrn_b200.engine import Engine
    from kcrn_b200.synth import make_problem_device
    from oracle import oracle as O
    n, d = 100_000, 10_000
    indptr, indices, data, b = make_problem_device(n, d, 116.0, seed=1000, perm_seed=7)
    eng = Engine(indptr, indices, data, b, (n, d), l2=0.0)
    ip, ix, dv, bh = (t.cpu().numpy() for t in (indptr, indices, data, b))
    L = O.OracleLogistic(ip, ix, dv, (n, d), bh, l2=0.0, threads=4)
    rng = np.random.default_rng(5)
    x = 0.3 * rng.standard_normal(d)
    v = rng.standard_normal(d)
    xd, vd = (torch.as_tensor(a, device="cuda") for a in (x, v))
    eng.set_x(xd)
    rel = lambda p, q: float(np.linalg.norm(p - q) / np.linalg.norm(q))
    f = L.value(x)
    assert abs(eng.value() - f) <= 1e-12 * abs(f)
    assert rel(eng.gradient().cpu().numpy(), L.gradient(x)) <= 1e-12
    assert rel(eng.hvp(vd).cpu().numpy(), L.hess_vec_prod(x, v)) <= 1e-12
    cp, ci, cx = eng.csc()
    op, oi, ox = L.csc()
    assert np.array_equal(cp, op) and np.array_equal(ci, oi) and np.array_equal(cx, ox)
    eng.close()
    # two outer iterations of Krylov CRN with the configuration's m = 100 (cubic.py:265-309) against the oracle's loop
    import scipy.sparse as sp
    from optimizer.cubic import Cubic_Krylov_LS
    from optimizer.loss import LogisticRegression
    loss = LogisticRegression(sp.csr_matrix((dv, ix, ip), shape=(n, d)), bh, l1=0, l2=0, store_mat_vec_prod=True)
    opt = Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=100, tolerance=0, label="c5", tqdm=False)
    x0 = np.full(d, 0.5)
    opt.run(x0, it_max=2)
    ora = O.KrylovCRN(L, subspace_dim=100).run(x0, it_max=2)
    f_gpu = np.array([h["f"] for h in opt.history])
    f_cpu = np.array([h["f"] for h in ora.history])
    assert len(f_gpu) == len(f_cpu) == 2
    assert np.max(np.abs(f_gpu - f_cpu) / np.abs(f_cpu)) <= 1e-9, (f_gpu, f_cpu)
    assert [h["M"] for h in opt.history] == [h["M"] for h in ora.history]
