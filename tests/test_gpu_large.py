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
