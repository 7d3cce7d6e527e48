"""CPU: pins oracle/ (the C + numpy restatement) to outputs of the unmodified reference
(tests/golden/*.npz written by tests/golden/make_golden.py).  Tolerances are stated per check."""
import numpy as np
import pytest

from conftest import rel_err
from kcrn_b200.synth import make_problem
from oracle import oracle as O


def _loss(P, l2=0.0, threads=1):
    return O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b, l2=l2, threads=threads)


def test_scalar_primitives_bitexact(golden):
    g = golden("prims")
    # same libm, same formulas: the restatement reproduces scipy.special.expit / numba logsig bit for bit here
    assert np.array_equal(O.expit(g["x"]), g["expit"])
    assert np.array_equal(O.logsig(g["x"]), g["logsig"])
    for i in range(5):
        assert np.array_equal(O.map_labels(g[f"label_in{i}"]), g[f"label_out{i}"])
    with pytest.raises(ValueError):
        O.map_labels(np.array([0., 1., 2.]))


def test_csr_tocsc_bitexact_with_duplicates(golden):
    g = golden("oracle_methods")
    n, d = g["dup_shape"]
    cp, ci, cx = O.csr_tocsc(int(n), int(d), g["dup_indptr"], g["dup_indices"], g["dup_data"])
    assert np.array_equal(cp, g["dup_csc_indptr"])
    assert np.array_equal(ci, g["dup_csc_indices"])
    assert np.array_equal(cx, g["dup_csc_data"])
    L = O.OracleLogistic(g["dup_indptr"], g["dup_indices"], g["dup_data"], (n, d), np.zeros(int(n)))
    assert np.array_equal(L.A_dot(g["dup_x"]), g["dup_Ax"])          # sequential sums == scipy csr_matvec
    assert np.array_equal(L.At_dot(g["dup_r"]), g["dup_ATr"])        # scatter order == scipy csc_matvec
    L.threads = 2
    assert np.array_equal(L.At_dot(g["dup_r"]), g["dup_ATr"])        # threaded gather keeps the order


@pytest.mark.parametrize("tag,name", [("tiny", "tiny"), ("tiny_l2", "tiny"), ("small", "small")])
def test_oracle_methods_match_reference(golden, tag, name):
    g = golden("oracle_methods")
    P = make_problem(name, seed=3)
    assert P.fingerprint() == str(g[f"{tag}_fp"]), "synthetic generator drifted from the fixture"
    cp, ci, cx = O.csr_tocsc(P.n, P.d, P.indptr, P.indices, P.data)
    assert np.array_equal(cp, g[f"{tag}_csc_indptr"]) and np.array_equal(ci, g[f"{tag}_csc_indices"])
    assert np.array_equal(cx, g[f"{tag}_csc_data"])
    for threads in (1, 3):
        L = _loss(P, float(g[f"{tag}_l2"]), threads)
        x, v, xs = g[f"{tag}_x"], g[f"{tag}_v"], g[f"{tag}_xsat"]
        assert np.array_equal(L.mat_vec_product(x), g[f"{tag}_Ax"])
        for xx, sfx in ((x, ""), (P.x0, "0"), (xs, "_sat")):
            # 1e-14: only the final mean / division rounding can differ from the reference
            assert abs(L.value(xx) - g[f"{tag}_value{sfx}"]) <= 1e-14 * abs(g[f"{tag}_value{sfx}"])
            assert rel_err(L.gradient(xx), g[f"{tag}_grad{sfx}"]) <= 1e-14
            assert rel_err(L.hess_vec_prod(xx, v), g[f"{tag}_hvp{sfx}"]) <= 1e-14
    if f"{tag}_hess" in g:
        L = _loss(P, float(g[f"{tag}_l2"]))
        I = g[f"{tag}_I"]
        assert rel_err(L.hessian(g[f"{tag}_x"]), g[f"{tag}_hess"]) <= 1e-13
        assert rel_err(L.partial_gradient(g[f"{tag}_x"], I), g[f"{tag}_pgrad"]) <= 1e-13
        assert rel_err(L.partial_hessian(g[f"{tag}_x"], I), g[f"{tag}_phess"]) <= 1e-13


@pytest.mark.parametrize("tag", ["m1", "m2", "m6", "m12", "brk_early_m8", "brk_last_m4", "brk_m3"])
def test_lanczos_matches_reference(golden, tag):
    g = golden("lanczos")
    S, g0, m = g[f"{tag}_S"], g[f"{tag}_g"], int(g[f"{tag}_m"])
    V, al, be, beta = O.lanczos(lambda v: S @ v, g0, m)
    assert V.shape == g[f"{tag}_V"].shape and al.shape == g[f"{tag}_alphas"].shape
    assert be.shape == g[f"{tag}_betas"].shape
    # numpy in both: identical operation sequence -> identical bits
    assert np.array_equal(V, g[f"{tag}_V"]) and np.array_equal(al, g[f"{tag}_alphas"])
    assert np.array_equal(be, g[f"{tag}_betas"]) and beta == float(g[f"{tag}_beta"])


def test_cubic_solver_matches_reference(golden):
    g = golden("cubic")
    worst = 0.0
    for i in range(int(g["count"])):
        al, be = g[f"c{i}_alphas"], g[f"c{i}_betas"]
        g0, M, r0, r_ref, dec_ref = g[f"c{i}_scal"]
        T = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)
        gv = np.zeros(al.size)
        gv[0] = g0
        s, its, r, dec = O.cubic_solver_root(gv, T, M, epsilon=1e-8, r0=r0)
        assert its == int(g[f"c{i}_its"]), f"case {i}: Newton iteration count"
        assert abs(r - r_ref) <= 1e-9 * abs(r_ref) + 1e-16
        assert abs(dec - dec_ref) <= 1e-9 * abs(dec_ref) + 1e-300
        worst = max(worst, rel_err(s, g[f"c{i}_s"]))
    assert worst <= 1e-8      # own Cholesky vs LAPACK dposv on ill-conditioned T + rI (cond up to ~1e9)
    Md, r0d, r_ref, dec_ref = g["dense_scal"]
    s, its, r, dec = O.cubic_solver_root(g["dense_g"], g["dense_H"], Md, r0=r0d)
    assert its == int(g["dense_its"]) and rel_err(s, g["dense_s"]) <= 1e-12 and abs(r - r_ref) <= 1e-12 * r_ref


@pytest.mark.parametrize("name,m,iters", [("small", 10, 30), ("w8a", 10, 8)])
def test_oracle_trajectory_tracks_reference(golden, name, m, iters):
    """Free-running oracle vs reference.  Identical arithmetic except libm-level rounding inside the
    C sweeps, so the prefix agrees far below the 1e-8 bar; see DESIGN.md for the chaos envelope."""
    g = golden(f"traj_{name}")
    P = make_problem(name, seed=0)
    assert P.fingerprint() == str(g["fp"])
    opt = O.KrylovCRN(_loss(P), subspace_dim=m).run(P.x0, it_max=iters)
    f = np.array([h["f"] for h in opt.history])
    ref = g[f"m{m}_ref_f"][:len(f)]
    # reference vs row-reversed reference = the reference's own rounding-noise envelope (running max)
    envelope = np.maximum.accumulate(np.abs(g[f"m{m}_refrev_f"][:len(f)] - ref) / np.abs(ref))
    dev = np.abs(f - ref) / np.abs(ref)
    assert np.all(dev <= np.maximum(1e-8, 10 * envelope)), (dev, envelope)
    assert np.array_equal(np.array([h["m_eff"] for h in opt.history]), g[f"m{m}_ref_m_eff"][:len(f)])
