"""CPU oracle for the Krylov-CRN hot path -- TEST INFRASTRUCTURE ONLY.

A float64 numpy + C (``kcrn_oracle.c``) restatement of what the reference computes on this path.
Nothing in the shipped package imports this module; it is used by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs as the
checker and the timed CPU baseline.

Reference functions restated (file:line in /root/reference):
  * ``OracleLogistic``      <- optimizer/loss.py:179-302  (labels :190-207, value :215-220,
                               gradient :223-232, partial_* :234-264, hessian :249-255,
                               Ax memo :266-286, hess_vec_prod :289-302, f_opt tracking :66-73)
  * ``lanczos``             <- optimizer/cubic.py:77-111  (no re-orthogonalisation)
  * ``lanczos(reorth=True)``   same recurrence + two classical Gram-Schmidt passes after the
                               ``w - alpha v`` update (SURVEY.md section 8(c) "reorth oracle"; not in the reference)
  * ``newton_scalar``       <- scipy/optimize/_zeros_py.py:319-358 (the loop root_scalar(method='newton') runs)
  * ``cubic_solver_root``   <- optimizer/cubic.py:40-75
  * ``KrylovCRN``           <- optimizer/cubic.py:238-319 + optimizer/optimizer.py:63-162 (loop, tolerance test)

Parity status: pinned against outputs of the unmodified reference executed in the build container
(tests/golden/*.npz, written by tests/golden/make_golden.py).  The reference itself has no tests.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    """Compile libkcrn_oracle.so next to this file (gcc; OpenMP when available)."""
    so = os.path.join(_HERE, "libkcrn_oracle.so")
    src = os.path.join(_HERE, "kcrn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        base = ["-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src, "-lm"]
        last = None
        for cc, extra in (("/usr/bin/gcc", ["-fopenmp"]), ("gcc", ["-fopenmp"]), ("gcc", [])):
            try:
                subprocess.run([cc] + extra + base, check=True, capture_output=True)
                last = None
                break
            except (OSError, subprocess.CalledProcessError) as e:  # try the next compiler
                last = e
        if last is not None:
            raise RuntimeError(f"could not build the CPU oracle: {last}")
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.ko_max_threads.restype = ctypes.c_int
        L.ko_csr_tocsc.argtypes = [_i64, _i64, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        L.ko_csr_matvec.argtypes = [_i64, _i32p, _i32p, _f64p, _f64p, _f64p, ctypes.c_int]
        L.ko_csc_matvec.argtypes = [_i64, _i64, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.ko_cscT_gather.argtypes = [_i64, _i32p, _i32p, _f64p, _f64p, _f64p, ctypes.c_int]
        L.ko_expit.argtypes = [_i64, _f64p, _f64p]
        L.ko_logsig.argtypes = [_i64, _f64p, _f64p]
        L.ko_hvp_weights.argtypes = [_i64, _f64p, _f64p]
        L.ko_posv.argtypes = [_i64, _f64p, _f64p, _f64p, _f64p]
        L.ko_posv.restype = ctypes.c_int
        _LIB = L
    return _LIB


class NotPositiveDefinite(np.linalg.LinAlgError):
    """Raised where LAPACK dposv would report info > 0 (scipy.linalg.solve(assume_a='pos'))."""


# ----------------------------------------------------------------------------- primitives
def csr_tocsc(n, d, indptr, indices, data):
    cp = np.empty(d + 1, np.int32)
    ci = np.empty(indices.size, np.int32)
    cx = np.empty(indices.size, np.float64)
    lib().ko_csr_tocsc(n, d, indptr, indices, data, cp, ci, cx)
    return cp, ci, cx


def expit(x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    lib().ko_expit(x.size, x, out)
    return out


def logsig(x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    lib().ko_logsig(x.size, x, out)
    return out


def posv(S, rhs):
    S = np.ascontiguousarray(S, np.float64)
    rhs = np.ascontiguousarray(rhs, np.float64)
    m = rhs.size
    z = np.empty(m)
    info = lib().ko_posv(m, S, rhs, z, np.empty(m * m))
    if info:
        raise NotPositiveDefinite(f"leading minor of order {info} is not positive definite")
    return z


def map_labels(b):
    """{0,1} kept, {1,2} -> b-1, {-1,1} -> (b+1)/2, any other pair -> (b == b[0])  (loss.py:190-207)."""
    b = np.asarray(b)
    u = np.unique(b)
    if u.size > 2:
        raise ValueError("The number of classes must be no more than 2 for binary classification.")
    if u.size == 2 and not np.array_equal(u, [0, 1]):
        if np.array_equal(u, [1, 2]):
            return (b - 1).astype(np.float64)
        if np.array_equal(u, [-1, 1]):
            return ((b + 1) / 2).astype(np.float64)
        return 1.0 * (b == b[0])
    return b.astype(np.float64)


# ----------------------------------------------------------------------------- the loss oracle
class OracleLogistic:
    """f(x) = mean((1-b) Ax - logsig(Ax)) + l2/2 |x|^2 on a CSR matrix, as the reference evaluates it."""

    def __init__(self, indptr, indices, data, shape, b, l2=0.0, threads=1):
        self.n, self.dim = int(shape[0]), int(shape[1])
        self.indptr = np.ascontiguousarray(indptr, np.int32)
        self.indices = np.ascontiguousarray(indices, np.int32)
        self.data = np.ascontiguousarray(data, np.float64)
        self.b = map_labels(b)
        self.l2 = float(l2)
        self.threads = int(threads)
        self._csc = None
        self.f_opt, self.x_opt = np.inf, None
        self.reset()
        self.calls = {"value": 0, "gradient": 0, "hvp": 0, "sweeps": 0}

    # -- cache (loss.py:266-286): value-equality memo of A@x
    def reset(self):
        self._x_last, self._Ax = None, None

    def csc(self):
        if self._csc is None:
            self._csc = csr_tocsc(self.n, self.dim, self.indptr, self.indices, self.data)
        return self._csc

    def A_dot(self, x):
        y = np.empty(self.n)
        lib().ko_csr_matvec(self.n, self.indptr, self.indices, self.data,
                            np.ascontiguousarray(x, np.float64), y, self.threads)
        self.calls["sweeps"] += 1
        return y

    def At_dot(self, r):
        y = np.empty(self.dim)
        r = np.ascontiguousarray(r, np.float64)
        if self.threads > 1:
            cp, ci, cx = self.csc()
            lib().ko_cscT_gather(self.dim, cp, ci, cx, r, y, self.threads)
        else:
            lib().ko_csc_matvec(self.n, self.dim, self.indptr, self.indices, self.data, r, y)
        self.calls["sweeps"] += 1
        return y

    def mat_vec_product(self, x):
        if self._x_last is not None and np.array_equal(x, self._x_last):
            return self._Ax
        self._Ax = self.A_dot(x)
        self._x_last = np.array(x, copy=True)
        return self._Ax

    def value(self, x):
        self.calls["value"] += 1
        Ax = self.mat_vec_product(x)
        f = np.mean((1 - self.b) * Ax - logsig(Ax))
        if self.l2 != 0:
            f = f + self.l2 / 2 * np.linalg.norm(x) ** 2
        if f < self.f_opt:
            self.f_opt, self.x_opt = f, np.array(x, copy=True)
        return f

    def gradient(self, x):
        self.calls["gradient"] += 1
        resid = expit(self.mat_vec_product(x)) - self.b
        g = self.At_dot(resid) / self.n
        return g if self.l2 == 0 else g + self.l2 * x

    def hess_weights(self, x):
        w = np.empty(self.n)
        lib().ko_hvp_weights(self.n, np.ascontiguousarray(self.mat_vec_product(x)), w)
        return w

    def hess_vec_prod(self, x, v):
        self.calls["hvp"] += 1
        t = self.hess_weights(x) * self.A_dot(v)
        return self.At_dot(t) / self.n + self.l2 * v

    # -- config-4 pieces (dense; only for small dim)
    def dense(self):
        M = np.zeros((self.n, self.dim))
        rows = np.repeat(np.arange(self.n), np.diff(self.indptr))
        np.add.at(M, (rows, self.indices), self.data)
        return M

    def hessian(self, x):
        Ad = self.dense()
        return Ad.T @ (self.hess_weights(x)[:, None] * Ad) / self.n + self.l2 * np.eye(self.dim)

    def partial_gradient(self, x, I):
        resid = expit(self.mat_vec_product(x)) - self.b
        g = self.At_dot(resid)[I] / self.n
        return g if self.l2 == 0 else g + self.l2 * x[I]

    def partial_hessian(self, x, I):
        AI = self.dense()[:, I]
        return AI.T @ (self.hess_weights(x)[:, None] * AI) / self.n + self.l2 * np.eye(len(I))

    @staticmethod
    def norm(x):
        return np.linalg.norm(x)


# ----------------------------------------------------------------------------- Lanczos
def lanczos(op, g, m=10, reorth=False):
    """m-step Lanczos on the symmetric operator ``op`` started from g/|g|.

    Returns (V[d, m_eff], alphas[m_eff], betas[m_eff-1], beta_last).  Breakdown is the absolute
    test beta < 1e-6; on breakdown before the last step the outputs are cut to the columns built
    so far; the last diagonal entry is always recomputed as v.op(v) with the current v.
    """
    d = g.shape[0]
    beta = 0.0
    prev = np.zeros_like(g)
    v = g / np.linalg.norm(g)
    V = np.zeros((d, m))
    V[:, 0] = v
    alphas, betas = np.zeros(m), np.zeros(m - 1)
    stop = m - 1                       # index of the last column that holds a vector
    for j in range(m - 1):
        w = op(v) - beta * prev
        a = np.dot(v, w)
        alphas[j] = a
        w = w - a * v
        if reorth:
            Q = V[:, :j + 1]
            for _ in range(2):
                w = w - Q @ (Q.T @ w)
        beta = np.linalg.norm(w)
        if abs(beta) < 1e-6:
            stop = j
            break
        betas[j] = beta
        prev, v = v, w / beta
        V[:, j + 1] = v
    if m > 1 and stop < m - 2:
        V, alphas, betas = V[:, :stop + 1], alphas[:stop + 1], betas[:stop]
    alphas[-1] = np.dot(v, op(v))
    return V, alphas, betas, beta


# ----------------------------------------------------------------------------- sub-problem
def newton_scalar(f, fprime, x0, tol, maxiter):
    """Newton-Raphson exactly as scipy's scalar ``newton`` iterates (disp=False, rtol=0)."""
    p0 = float(x0)
    p = p0
    for itr in range(maxiter):
        fv = f(p0)
        if fv == 0:
            return p0, itr, True
        fd = fprime(p0)
        if fd == 0:
            return p0, itr + 1, False
        p = p0 - fv / fd
        if abs(p - p0) <= tol:
            return p, itr + 1, True
        p0 = p
    return p, maxiter, False


def cubic_solver_root(g, H, M, it_max=100, epsilon=1e-8, r0=0.1):
    """argmin_s <g,s> + 1/2 <s,Hs> + M/3 |s|^3 through the secular equation r^2 = M^2 |s_r|^2,
    s_r = -(H + r I)^-1 g.  Returns (s, newton_iterations, r, model_decrease)."""
    g = np.asarray(g, np.float64)
    H = np.asarray(H, np.float64)
    eye = np.eye(g.size)

    def s_of(lam):
        return -posv(H + lam * eye, g)

    def phi(lam):
        return lam ** 2 - M ** 2 * np.linalg.norm(s_of(lam)) ** 2

    def dphi(lam):
        s = s_of(lam)
        return 2 * lam + 2 * M ** 2 * np.dot(s, posv(H + lam * eye, s))

    r, its, _ = newton_scalar(phi, dphi, r0, epsilon, it_max)
    s = s_of(r)
    ns = np.linalg.norm(s)
    return s, its, r, r / 2 * ns ** 2 - M / 3 * ns ** 3 - np.dot(g, s) / 2


# ----------------------------------------------------------------------------- outer loop
class KrylovCRN:
    """Krylov cubic-regularised Newton with backtracking on M (cubic.py:238-319), plus the
    run loop's iteration / time / |dx| < tolerance stopping rules (optimizer.py:104-113)."""

    def __init__(self, loss, reg_coef=1e-3, subspace_dim=10, solver_eps=1e-8, beta=0.5,
                 tolerance=1e-9, reorth=False):
        self.loss, self.M, self.m = loss, float(reg_coef), int(subspace_dim)
        self.eps, self.shrink, self.tol, self.reorth = solver_eps, beta, tolerance, reorth
        self.r0, self.value, self.solver_it = 0.1, None, 0
        self.history = []   # per outer iteration: dict(f, M, solver_it, m_eff, backtracks)

    def step(self):
        L = self.loss
        if self.value is None:
            self.value = L.value(self.x)
        g = L.gradient(self.x)
        x = self.x
        V, al, be, _ = lanczos(lambda v: L.hess_vec_prod(x, v), g, self.m, self.reorth)
        T = np.diag(al) + np.diag(be, -1) + np.diag(be, 1)
        gs = np.zeros(al.size)
        gs[0] = np.linalg.norm(g)
        M = self.M * self.shrink
        s, its, r, dec = cubic_solver_root(gs, T, M, epsilon=self.eps, r0=self.r0)
        xn = x + V @ s
        fn = L.value(xn)
        tries = 0
        while fn > self.value - dec and tries < 20:
            M = M / self.shrink
            s, its, r, dec = cubic_solver_root(gs, T, M, epsilon=self.eps, r0=self.r0)
            xn = x + V @ s
            fn = L.value(xn)
            tries += 1
        self.x, self.M, self.value, self.r0 = xn, M, fn, r
        self.solver_it += its
        self.history.append(dict(f=fn, M=M, solver_it=self.solver_it, m_eff=al.size,
                                 backtracks=tries, r=r))

    def run(self, x0, it_max=100, t_max=np.inf):
        self.x = np.array(x0, copy=True)
        self.loss.reset()
        self.xs, self.ts = [self.x.copy()], [0.0]
        t0 = time.perf_counter()
        it, x_old = 0, None
        while it < it_max and time.perf_counter() - t0 < t_max:
            if self.tol > 0 and x_old is not None and np.linalg.norm(self.x - x_old) < self.tol:
                break
            x_old = self.x.copy()
            self.step()
            it += 1
            self.xs.append(self.x.copy())
            self.ts.append(time.perf_counter() - t0)
        return self
