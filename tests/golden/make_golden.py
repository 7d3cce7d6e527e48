#!/usr/bin/env python
"""Writes tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this container.

Run once, here (the GPU box has no /root/reference):   python tests/golden/make_golden.py
Shims, exactly as SURVEY.md section 8(c): a stub ``matplotlib`` module (opt_trace.py:12 imports pyplot at
module level; matplotlib is not installed).  No reference source is edited or copied.

Inputs are regenerated on the test side from kcrn_b200.synth (same seeds); each fixture stores
the input fingerprint so a drifting generator is detected instead of silently compared.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# only the synthetic generator comes from this repository (loaded by file path: the package directory also
# holds OUR ``optimizer`` mirror, which must not shadow the reference's ``optimizer`` package imported below)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "kcrn_synth", os.path.join(ROOT, "krylov-cubic-regularized-newton_b200", "kcrn_b200", "synth.py"))
_synth = importlib.util.module_from_spec(_spec)
sys.modules["kcrn_synth"] = _synth
_spec.loader.exec_module(_synth)

_mpl, _plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
_mpl.pyplot = _plt
sys.modules.update({"matplotlib": _mpl, "matplotlib.pyplot": _plt})
sys.path.insert(0, "/root/reference")

import scipy.sparse as sp  # noqa: E402
import optimizer.cubic as ref_cubic  # noqa: E402  (the reference's package)
from optimizer.loss import LogisticRegression, logsig  # noqa: E402
import scipy.special  # noqa: E402

make_problem = _synth.make_problem

assert ref_cubic.__file__.startswith("/root/reference"), ref_cubic.__file__
warnings.filterwarnings("ignore")
_ref_lanczos = ref_cubic.Lanczos


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ------------------------------------------------------------------ 1. scalar primitives
def prims():
    edge = np.array([-800., -745.2, -40., -33.0000001, -33., -32.9999999, -18.0000001, -18., -17.9999999,
                     -1., -1e-9, 0., 1e-9, 1., 17., 36.5, 36.999, 37., 37.0000001, 40., 709., 800.])
    rng = np.random.default_rng(11)
    xs = np.concatenate([edge, rng.uniform(-45, 45, 500), rng.standard_normal(500)])
    label_in = [np.array([-1., 1., 1., -1.]), np.array([1., 2., 2., 1.]), np.array([0., 1., 1., 0.]),
                np.array([3., 5., 5., 3., 3.]), np.array([5., 3., 3.])]
    label_out = []
    Adummy = sp.csr_matrix(np.eye(5))
    for b in label_in:
        label_out.append(np.asarray(LogisticRegression(Adummy[:len(b)], b).b, dtype=np.float64))
    save("prims", x=xs, logsig=logsig(xs.copy()), expit=scipy.special.expit(xs),
         **{f"label_in{i}": a for i, a in enumerate(label_in)},
         **{f"label_out{i}": a for i, a in enumerate(label_out)})


# ------------------------------------------------------------------ 2. oracle methods on small matrices
def oracle_methods():
    out = {}
    for tag, name, l2 in (("tiny", "tiny", 0.0), ("tiny_l2", "tiny", 0.05), ("small", "small", 0.0)):
        P = make_problem(name, seed=3)
        A = P.scipy_csr()
        loss = LogisticRegression(A, P.b, l1=0, l2=l2, store_mat_vec_prod=True)
        rng = np.random.default_rng(5)
        x = rng.standard_normal(P.d)
        v = rng.standard_normal(P.d)
        xsat = 400.0 * rng.standard_normal(P.d)          # drives |Ax| far past the +-37 switches
        C = A.tocsc()
        out.update({
            f"{tag}_fp": np.array(P.fingerprint()), f"{tag}_l2": np.array(l2),
            f"{tag}_x": x, f"{tag}_v": v, f"{tag}_xsat": xsat,
            f"{tag}_csc_indptr": C.indptr.astype(np.int32), f"{tag}_csc_indices": C.indices.astype(np.int32),
            f"{tag}_csc_data": C.data,
            f"{tag}_value": np.array(loss.value(x)), f"{tag}_grad": loss.gradient(x),
            f"{tag}_hvp": loss.hess_vec_prod(x, v), f"{tag}_Ax": loss.mat_vec_product(x).copy(),
            f"{tag}_value0": np.array(loss.value(P.x0)), f"{tag}_grad0": loss.gradient(P.x0),
            f"{tag}_hvp0": loss.hess_vec_prod(P.x0, v),
            f"{tag}_value_sat": np.array(loss.value(xsat)), f"{tag}_grad_sat": loss.gradient(xsat),
            f"{tag}_hvp_sat": loss.hess_vec_prod(xsat, v),
        })
        if P.d <= 100:
            I = np.random.default_rng(42).choice(P.d, size=7, replace=False)
            out[f"{tag}_hess"] = np.asarray(loss.hessian(x))
            out[f"{tag}_I"] = I.astype(np.int64)
            out[f"{tag}_pgrad"] = np.asarray(loss.partial_gradient(x, I)).ravel()
            out[f"{tag}_phess"] = np.asarray(loss.partial_hessian(x, I).todense())
            delta = rng.standard_normal(7)
            Ax = loss.mat_vec_product(x).copy()
            loss.update_mat_vec_product(Ax, delta, I)
            out[f"{tag}_delta"] = delta
            out[f"{tag}_Ax_upd"] = np.asarray(loss._mat_vec_prod).ravel()

    # CSR -> CSC on a matrix with unsorted columns, duplicate entries, empty rows and empty columns
    rng = np.random.default_rng(8)
    n, d = 40, 23
    lens = rng.integers(0, 9, n)
    lens[[3, 17]] = 0
    indptr = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
    indices = rng.integers(0, d - 3, indptr[-1]).astype(np.int32)   # last 3 columns empty; duplicates likely
    data = rng.standard_normal(indptr[-1])
    M = sp.csr_matrix((data, indices, indptr), shape=(n, d))
    C = M.tocsc()
    r = rng.standard_normal(n)
    x = rng.standard_normal(d)
    out.update(dup_indptr=indptr, dup_indices=indices, dup_data=data, dup_shape=np.array([n, d]),
               dup_csc_indptr=C.indptr.astype(np.int32), dup_csc_indices=C.indices.astype(np.int32),
               dup_csc_data=C.data, dup_r=r, dup_x=x, dup_Ax=M @ x, dup_ATr=M.T @ r)
    save("oracle_methods", **out)


# ------------------------------------------------------------------ 3. Lanczos (cubic.py:77-111)
def lanczos_cases():
    out = {}
    rng = np.random.default_rng(21)
    d = 48
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    S = (Q * np.linspace(0.5, 3.0, d)) @ Q.T            # well conditioned SPD: stays orthogonal for small m
    S = (S + S.T) / 2
    g = rng.standard_normal(d)
    cases = {"m1": (S, g, 1), "m2": (S, g, 2), "m6": (S, g, 6), "m12": (S, g, 12)}
    # breakdown: start vector inside a 3-dimensional invariant subspace -> beta ~ 0 at j = 2
    lam = np.concatenate(([1.0, 2.0, 3.5], np.linspace(4, 9, d - 3)))
    S3 = (Q * lam) @ Q.T
    S3 = (S3 + S3.T) / 2
    g3 = Q[:, :3] @ np.array([1.0, -2.0, 0.7])
    cases["brk_early_m8"] = (S3, g3, 8)     # break at j=2 <  m-2 -> truncated to 3 columns
    cases["brk_last_m4"] = (S3, g3, 4)      # break at j=2 == m-2 -> not truncated, zero last column
    cases["brk_m3"] = (S3, g3, 3)           # m-1 = 2 loop steps, no break inside the loop
    for tag, (S_, g_, m) in cases.items():
        V, al, be, beta = _ref_lanczos(lambda v, S_=S_: S_ @ v, g_.copy(), m=m)
        out.update({f"{tag}_S": S_, f"{tag}_g": g_, f"{tag}_m": np.array(m), f"{tag}_V": V,
                    f"{tag}_alphas": al, f"{tag}_betas": be, f"{tag}_beta": np.array(beta)})
    save("lanczos", **out)


# ------------------------------------------------------------------ 4. sub-problem (cubic.py:40-75)
def cubic_cases():
    rng = np.random.default_rng(33)
    recs = []
    for m in (1, 2, 3, 5, 10, 20, 50, 100):
        for rep in range(6):
            al = rng.uniform(0.02, 2.0, m) * 10.0 ** rng.uniform(-3, 0)
            be = rng.uniform(0.0, 0.4, max(m - 1, 0)) * al.min()
            T = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)
            T += (max(0.0, -np.linalg.eigvalsh(T).min()) + 10.0 ** rng.uniform(-9, -2)) * np.eye(m)
            g = np.zeros(m)
            g[0] = 10.0 ** rng.uniform(-6, 1)
            M = 1e-3 * 2.0 ** rng.integers(-30, 8)
            r0 = [0.1, 1e-3, 5.0][rep % 3]
            s, its, r, dec = ref_cubic.cubic_solver_root(g, T, M, epsilon=1e-8, r0=r0)
            recs.append((m, np.diag(T).copy(), np.diag(T, 1).copy(), g[0], M, r0, s, its, r, dec))
    out = {"count": np.array(len(recs))}
    for i, (m, a, b_, g0, M, r0, s, its, r, dec) in enumerate(recs):
        out.update({f"c{i}_alphas": a, f"c{i}_betas": b_, f"c{i}_scal": np.array([g0, M, r0, r, dec]),
                    f"c{i}_its": np.array(its), f"c{i}_s": s})
    # dense (non-tridiagonal) SPD case, as SSCN / full CRN would pass
    B = rng.standard_normal((12, 12))
    H = B @ B.T / 12 + 0.05 * np.eye(12)
    gd = rng.standard_normal(12)
    s, its, r, dec = ref_cubic.cubic_solver_root(gd, H, 0.37, epsilon=1e-8, r0=0.1)
    out.update(dense_H=H, dense_g=gd, dense_scal=np.array([0.37, 0.1, r, dec]), dense_its=np.array(its), dense_s=s)
    save("cubic", **out)


# ------------------------------------------------------------------ 5. trajectories (cubic.py:238-319)
def lanczos_reorth(A, v, m=10):
    """SURVEY.md section 8(c) 'reorth oracle': the reference recurrence plus two classical Gram-Schmidt
    passes after the alpha update.  Everything else in the step stays the reference's code."""
    beta = 0
    v_pre = np.zeros_like(v)
    v = v / np.linalg.norm(v)
    V = np.zeros((len(v), m))
    V[:, 0] = v
    alphas, betas = np.zeros(m), np.zeros(m - 1)
    j = -1
    for j in range(m - 1):
        w = A(v) - beta * v_pre
        alpha = np.dot(v, w)
        alphas[j] = alpha
        w = w - alpha * v
        Q = V[:, :j + 1]
        for _ in range(2):
            w = w - Q @ (Q.T @ w)
        beta = np.linalg.norm(w)
        if np.abs(beta) < 1e-6:
            break
        betas[j] = beta
        v_pre = v
        v = w / beta
        V[:, j + 1] = v
    if m > 1 and j < m - 2:
        V, alphas, betas = V[:, :j + 1], alphas[:j + 1], betas[:j]
    alphas[-1] = np.dot(v, A(v))
    return V, alphas, betas, beta


def run_reference(P, m, iters, reorth=False, reverse=False, keep_x=()):
    A = P.scipy_csr()
    b = P.b
    if reverse:
        A, b = A[::-1].tocsr(), b[::-1].copy()
    loss = LogisticRegression(A, b, l1=0, l2=0, store_mat_vec_prod=True)
    m_effs = []

    def spy(op, v, m=10):
        res = (lanczos_reorth if reorth else _ref_lanczos)(op, v, m=m)
        m_effs.append(len(res[1]))
        return res

    ref_cubic.Lanczos = spy
    opt = ref_cubic.Cubic_Krylov_LS(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, label="g")
    opt.init_run(P.x0)
    rec = dict(f=[], M=[], solver_it=[], r0=[], dx=[])
    xs = {}
    for k in range(iters):
        if k in keep_x:
            xs[k] = (opt.x.copy(), opt.value, opt.reg_coef, opt.r0)
        x_old = opt.x.copy()
        opt.step()
        rec["f"].append(opt.value); rec["M"].append(opt.reg_coef)
        rec["solver_it"].append(opt.solver_it); rec["r0"].append(opt.r0)
        rec["dx"].append(np.linalg.norm(opt.x - x_old))
        if rec["dx"][-1] < 1e-9:
            break
    ref_cubic.Lanczos = _ref_lanczos
    out = {k: np.array(v, dtype=np.float64) for k, v in rec.items()}
    out["m_eff"] = np.array(m_effs, dtype=np.int64)
    out["x_final"] = opt.x.copy()
    out["f_opt"] = np.array(loss.f_opt)
    return out, xs


def trajectories():
    plan = [("small", (10,), 30, (0, 3, 8)), ("w8a", (10,), 30, (0, 1, 2, 5, 12)), ("rcv1", (10, 50), 40, (0, 6))]
    for name, ms, iters, keep in plan:
        P = make_problem(name, seed=0)
        out = {"fp": np.array(P.fingerprint()), "nnz": np.array(P.nnz)}
        for m in ms:
            base, xs = run_reference(P, m, iters, keep_x=keep)
            variants = {"ref": base,
                        "refrev": run_reference(P, m, iters, reverse=True)[0],
                        "ro": run_reference(P, m, iters, reorth=True)[0],
                        "rorev": run_reference(P, m, iters, reorth=True, reverse=True)[0]}
            for vn, rec in variants.items():
                for k, arr in rec.items():
                    if k == "x_final" and (vn != "ref" or P.d > 1000):
                        continue
                    out[f"m{m}_{vn}_{k}"] = arr
            for k, (x, val, M, r0) in xs.items():
                if P.d > 1000 and k > 0 and m != ms[0]:
                    continue
                out[f"m{m}_state{k}_x"] = x
                out[f"m{m}_state{k}_scal"] = np.array([val if val is not None else np.nan, M, r0])
            # teacher-forced single steps from the stored states: run one reference step from each
            for k, (x, val, M, r0) in xs.items():
                if f"m{m}_state{k}_x" not in out:
                    continue
                loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
                opt = ref_cubic.Cubic_Krylov_LS(loss=loss, reg_coef=M, subspace_dim=m, tolerance=1e-9, label="g")
                opt.init_run(x)
                opt.value, opt.r0 = val, r0
                opt.step()
                if P.d <= 1000:
                    out[f"m{m}_state{k}_xnext"] = opt.x.copy()
                else:
                    out[f"m{m}_state{k}_xnext_norm"] = np.array([np.linalg.norm(opt.x), np.linalg.norm(opt.x - x)])
                out[f"m{m}_state{k}_next"] = np.array([opt.value, opt.reg_coef, opt.r0, opt.solver_it])
            print(name, m, "ref f[-1]=%.6e" % base["f"][-1],
                  "self-noise max rel=%.2e" % np.max(np.abs(base["f"][:len(variants['refrev']['f'])] - variants["refrev"]["f"][:len(base['f'])]) / np.abs(base["f"][:len(variants['refrev']['f'])])))
        save(f"traj_{name}", **out)


# ------------------------------------------------------------------ 5b. C3 (news20-shape, m = 50) and f*
def formula_point(d, k):
    """A deterministic non-trivial iterate both sides can rebuild (d = 1.36 M is too large to store)."""
    i = np.arange(d, dtype=np.float64)
    return 0.3 * np.cos(1e-3 * i * (k + 1)) + 0.05 * np.sin(0.37 * i)


def news20():
    """BASELINE configs[2]: teacher-forced single steps (from x0 and from two formula points) and a short free-running
    trajectory with its self-noise envelope, all from the unmodified reference."""
    P = make_problem("news20", seed=0)
    m, iters = 50, 8
    out = {"fp": np.array(P.fingerprint()), "nnz": np.array(P.nnz)}
    variants = {"ref": run_reference(P, m, iters)[0],
                "refrev": run_reference(P, m, iters, reverse=True)[0],
                "ro": run_reference(P, m, iters, reorth=True)[0],
                "rorev": run_reference(P, m, iters, reorth=True, reverse=True)[0]}
    for vn, rec in variants.items():
        for k, arr in rec.items():
            if k != "x_final":
                out[f"m{m}_{vn}_{k}"] = arr
    states = {0: (P.x0, None, 1e-3, 0.1), 1: (formula_point(P.d, 1), None, 1e-5, 0.05),
              2: (formula_point(P.d, 2), None, 4e-7, 0.01)}
    for k, (x, val, M, r0) in states.items():
        loss = LogisticRegression(P.scipy_csr(), P.b, l1=0, l2=0, store_mat_vec_prod=True)
        opt = ref_cubic.Cubic_Krylov_LS(loss=loss, reg_coef=M, subspace_dim=m, tolerance=1e-9, label="g")
        opt.init_run(x)
        opt.value, opt.r0 = val, r0
        opt.step()
        out[f"m{m}_state{k}_scal"] = np.array([np.nan, M, r0])
        out[f"m{m}_state{k}_xnext_norm"] = np.array([np.linalg.norm(opt.x), np.linalg.norm(opt.x - x)])
        out[f"m{m}_state{k}_next"] = np.array([opt.value, opt.reg_coef, opt.r0, opt.solver_it])
        print("news20 state", k, out[f"m{m}_state{k}_next"], out[f"m{m}_state{k}_xnext_norm"])
    save("traj_news20", **out)


def fstar():
    """f* per data set for the suboptimality f(x_k) - f* (SURVEY 8(d) step 3): the minimum over EVERYTHING the
    reference evaluated on that input (loss.py:70-72 keeps f_opt per oracle; cubic_newton.py:140 takes the minimum
    over all methods run) -- here every stored variant plus one long re-orthogonalised run, which goes lowest."""
    out = {}
    for name, m, iters in (("small", 10, 90), ("w8a", 10, 90), ("rcv1", 50, 100), ("news20", 50, 14)):
        P = make_problem(name, seed=0)
        rec, _ = run_reference(P, m, iters, reorth=True)
        best = float(rec["f_opt"])
        path = os.path.join(HERE, f"traj_{name}.npz")
        if os.path.exists(path):
            with np.load(path) as z:
                for k in z.files:
                    if k.endswith("_f_opt"):
                        best = min(best, float(z[k]))
        out[name] = np.array(best)
        out[name + "_long_f"] = rec["f"]
        print(name, "f* = %.12e after %d iterations" % (best, len(rec["f"])))
    save("fstar", **out)


# ------------------------------------------------------------------ 6. config-4 baselines (cubic.py:115-235, 321-408)
def config4():
    """SSCN (m = 10, 50) and full-Hessian CRN on the 'small' problem, plus the same runs on row-reversed data
    (the reference's own rounding-noise envelope)."""
    P = make_problem("small", seed=0)
    out = {"fp": np.array(P.fingerprint())}

    def run(kind, m, iters, reverse):
        A, b = P.scipy_csr(), P.b
        if reverse:
            A, b = A[::-1].tocsr(), b[::-1].copy()
        if kind == "sscn":
            loss = LogisticRegression(A.tocsc(), b, l1=0, l2=0, store_mat_vec_prod=True)
            opt = ref_cubic.SSCN(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, label="s")
        else:
            loss = LogisticRegression(A, b, l1=0, l2=0, store_mat_vec_prod=True)
            opt = ref_cubic.Cubic_LS(loss=loss, reg_coef=1e-3, cubic_solver="full" if kind == "full" else "CG",
                                     tolerance=1e-8, label="c")
        opt.rng = np.random.default_rng(42)
        opt.init_run(P.x0)
        f, M, sit = [], [], []
        for _ in range(iters):
            opt.step()
            f.append(opt.value); M.append(opt.reg_coef); sit.append(opt.solver_it)
        return np.array(f), np.array(M), np.array(sit, dtype=np.int64), opt.x.copy()

    # cg(..., tol=) of the pinned scipy 1.10.1 is cg(..., rtol=) in the installed one
    from scipy.sparse.linalg import cg as _scipy_cg
    ref_cubic.cg = lambda A, b, tol: _scipy_cg(A, b, rtol=tol)

    for kind, m, iters in (("sscn", 10, 40), ("sscn", 50, 40), ("full", 0, 12), ("cg", 0, 12)):
        tag = f"{kind}{m}" if kind == "sscn" else kind
        f, M, sit, x = run(kind, m, iters, False)
        fr, _, _, _ = run(kind, m, iters, True)
        out.update({f"{tag}_f": f, f"{tag}_M": M, f"{tag}_solver_it": sit, f"{tag}_x_final": x, f"{tag}_rev_f": fr})
        print(tag, "f[-1] = %.6e" % f[-1], "self-noise max rel = %.2e" % np.max(np.abs(f - fr) / np.abs(f)))
    save("traj_config4", **out)


def config4_l2():
    """SSCN with an l2 term (the value and the block gradient then depend on x itself, not only on A x: the cached-A x
    override of update_mat_vec_product must not freeze x -- advisor finding of round 1)."""
    P = make_problem("small", seed=0)
    out = {"fp": np.array(P.fingerprint()), "l2": np.array(1e-3)}
    for m in (10,):
        loss = LogisticRegression(P.scipy_csr().tocsc(), P.b, l1=0, l2=1e-3, store_mat_vec_prod=True)
        opt = ref_cubic.SSCN(loss=loss, reg_coef=1e-3, subspace_dim=m, tolerance=1e-9, label="s")
        opt.rng = np.random.default_rng(42)
        opt.init_run(P.x0)
        f, M, sit = [], [], []
        for _ in range(40):
            opt.step()
            f.append(opt.value); M.append(opt.reg_coef); sit.append(opt.solver_it)
        out.update({f"sscn{m}_f": np.array(f), f"sscn{m}_M": np.array(M), f"sscn{m}_solver_it": np.array(sit, dtype=np.int64),
                    f"sscn{m}_x_final": opt.x.copy()})
        print("sscn l2", m, "f[-1] = %.6e" % f[-1])
    save("traj_config4_l2", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["prims", "oracle_methods", "lanczos", "cubic", "traj", "config4", "config4_l2", "news20", "fstar"]
    if "config4" in which: config4()
    if "config4_l2" in which: config4_l2()
    if "prims" in which: prims()
    if "oracle_methods" in which: oracle_methods()
    if "lanczos" in which: lanczos_cases()
    if "cubic" in which: cubic_cases()
    if "traj" in which: trajectories()
    if "news20" in which: news20()
    if "fstar" in which: fstar()
