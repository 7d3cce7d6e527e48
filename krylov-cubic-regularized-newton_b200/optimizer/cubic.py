"""Krylov cubic-regularised Newton on the GPU (mirror of the reference's optimizer/cubic.py).

Public names and call signatures follow the reference:

    cubic_solver_root(g, H, M, it_max=100, epsilon=1e-8, r0=0.1) -> (s, iterations, r, model_decrease)   cubic.py:40-75
    Lanczos(A, v, m=10)                                          -> (V, alphas, betas, beta)             cubic.py:77-111
    Cubic_Krylov_LS(loss, reg_coef, subspace_dim, solver_eps, beta, ...)                                 cubic.py:238-319

``Cubic_Krylov_LS.step`` enqueues one outer iteration on the device -- gradient sweep, the whole m-step
Lanczos recurrence as one CUDA graph, the single-CTA sub-problem solve, the trial point, the forward
sweep and the value reduction -- and reads all of its scalars back with ONE device-to-host copy; the
backtracking loop on the regulariser M (cubic.py:294-303) stays on the host, as in the reference.
``reorth=True`` (not in the reference) switches the recurrence to full re-orthogonalisation.
"""
from __future__ import annotations

import time

import numpy as np

from kcrn_b200 import engine as _engine
from kcrn_b200._lib import KCRN_E_NOTPD
from optimizer.optimizer import Optimizer


def _is_tensor(x) -> bool:
    return hasattr(x, "data_ptr")


def _host(a):
    return a.detach().cpu().numpy() if _is_tensor(a) else np.asarray(a)


# --------------------------------------------------------------------------------- sub-problem
def cubic_solver_root(g, H, M, it_max=100, epsilon=1e-8, r0=0.1):
    """Solve  min_s <g, s> + 1/2 <s, H s> + M/3 |s|^3  by Newton's method on the secular equation.

    H must be (numerically) tridiagonal and g a multiple of e1 -- the form the Krylov method produces
    (cubic.py:278-283); that system is solved by the single-CTA kernel of libkcrn.  Dense blocks
    (SSCN, full CRN) go through ``optimizer.dense.cubic_solver_dense``.
    """
    import scipy.sparse as sp

    g_host = _host(g).astype(np.float64).ravel()
    m = g_host.size
    Hh = H.toarray() if sp.issparse(H) else _host(H)
    Hh = np.asarray(Hh, np.float64).reshape(m, m)
    off = Hh - np.diag(np.diag(Hh)) - np.diag(np.diag(Hh, 1), 1) - np.diag(np.diag(Hh, -1), -1)
    if np.any(off != 0.0) or np.any(g_host[1:] != 0.0) or not np.array_equal(np.diag(Hh, 1), np.diag(Hh, -1)):
        from optimizer.dense import cubic_solver_dense          # SSCN / full-CRN form: dense H, general g
        return cubic_solver_dense(g, Hh if not _is_tensor(H) else H, M, it_max=it_max, epsilon=epsilon, r0=r0)
    s, r, dec, its, status = _engine.cubic_tridiag(np.diag(Hh).copy(), np.diag(Hh, 1).copy(), g_host[0], M,
                                                   r0=r0, eps=epsilon, it_max=it_max)
    if status == KCRN_E_NOTPD:
        raise np.linalg.LinAlgError("H + r I is not positive definite")
    s = s if _is_tensor(g) else s.cpu().numpy()
    return s, its, r, dec


# --------------------------------------------------------------------------------- Lanczos
_workspaces = {}


def _workspace(d: int):
    """A matrix-free libkcrn handle (1 x d, no entries) that owns the basis and the recurrence state."""
    import torch

    key = (d, torch.cuda.current_device())
    if key not in _workspaces:
        _workspaces[key] = _engine.Engine(np.zeros(2, np.int32), np.zeros(0, np.int32), np.zeros(0), np.zeros(1),
                                          (1, d))
    return _workspaces[key]


def Lanczos(A, v, m=10, reorth=False):
    """m-step Lanczos for the symmetric operator ``A`` (a callable) started from v.

    The recurrence, its breakdown test (beta < 1e-6, absolute), the truncation on early breakdown and
    the final alpha follow cubic.py:77-111; the d-length arithmetic runs in libkcrn kernels.  Returns
    V (d x m_eff), alphas (m_eff), betas (m_eff - 1) and the last beta, as numpy arrays when v is numpy
    and as CUDA tensors otherwise.
    """
    import torch

    as_numpy = not _is_tensor(v)
    d = int(v.shape[0])
    ws = _workspace(d)
    g = torch.as_tensor(np.ascontiguousarray(v, np.float64), device=ws.device) if as_numpy else v.contiguous()
    ws.lanczos_begin(g, m, reorth)
    finished = False
    while not finished:
        cur = ws.lanczos_current()
        hv = A(cur.cpu().numpy()) if as_numpy else A(cur)
        hv = torch.as_tensor(np.ascontiguousarray(hv, np.float64), device=ws.device) if not _is_tensor(hv) else hv
        finished = ws.lanczos_feed(hv.contiguous())
    Vrows, alphas, bet, info = ws.lanczos_view()
    info_h = info.cpu().numpy()
    m_eff, nvalid = int(info_h[0]), int(info_h[1])
    V = Vrows[:m_eff].clone()
    if nvalid < m_eff:
        V[nvalid:] = 0.0            # breakdown at the last step: that column is never built (cubic.py:99)
    V = V.t()
    al = alphas[:m_eff].clone()
    be = bet[1:m_eff].clone()
    last_beta = float(bet[m].item())   # 4th return value of the reference: the last beta computed (cubic.py:111)
    if as_numpy:
        return np.ascontiguousarray(V.cpu().numpy()), al.cpu().numpy(), be.cpu().numpy(), last_beta
    return V, al, be, last_beta


# --------------------------------------------------------------------------------- the optimizer
class Cubic_Krylov_LS(Optimizer):
    """Krylov cubic regularized Newton method with line search.

    Arguments:
        reg_coef (float): estimate of the Hessian's Lipschitz constant (the regulariser M)
        subspace_dim (int): dimension m of the Krylov subspace
        solver_eps (float): accuracy of the Newton solve of the cubic sub-problem
        beta (float): backtracking factor for M
        reorth (bool): full re-orthogonalisation inside Lanczos (extension; default False = reference)
        device_step (bool, default False): step() runs one outer iteration INCLUDING its line search on the device
            (kcrn_krylov_run with one iteration: one graph launch and one synchronisation per step, however many
            trials it takes) instead of one graph launch and one host round trip per trial (kcrn_krylov_trial).  Same
            kernels in the same order, bit-identical results (test_device_step_equals_trial_step).  Measured on a
            B200 at the rcv1 shape it is SLOWER per step (0.36 vs 0.32 ms: a graph with WHILE nodes costs ~40 us more
            to launch than the two plain graphs cost in host round trips, profiles/r4c_bench_harness_ab.txt), so it is
            off by default; it pays when the host is slow or far away.  Row-sharded engines without the peer exchange
            always take the per-trial path.
        device_loop (bool): run() keeps the whole step loop on the device (kcrn_krylov_run: line search, commit and
            stopping rule in CUDA-graph WHILE nodes) in chunks of ``device_chunk`` iterations; same arithmetic, same
            trace, no per-iteration host round trip.  trace.ts then comes from the device clock.
    """

    def __init__(self, reg_coef=None, subspace_dim=100, solver_eps=1e-8, beta=0.5, reorth=False, device_loop=False,
                 device_chunk=64, device_step=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.device_loop = device_loop
        self.device_step = device_step
        self.device_chunk = device_chunk
        self.solver_it = 0
        self.subspace_dim = subspace_dim
        self.solver_eps = solver_eps
        self.beta = beta
        self.reorth = reorth
        self.r0 = 0.1
        self.value = None
        self.reg_coef = self.loss.hessian_lipschitz if reg_coef is None else reg_coef
        self.history = []       # per outer iteration: f, M, solver_it, m_eff, backtracks, r, hvps
        self._dx = None

    def step(self):
        loss, eng = self.loss, self.loss.engine
        if self.value is None:
            self.value = loss.value(self.x)
        else:
            loss._forward(self.x)                       # no-op when the caches already belong to x
        if self.device_step and (eng.world == 1 or eng.p2p):
            return self._step_on_device(loss, eng)
        M = self.reg_coef * self.beta                   # try a smaller M first                  cubic.py:286
        tries = 0
        first = True
        while True:
            # first trial: gradient (cubic.py:270), Lanczos (:273-275), sub-problem (:289-290), x + V s (:291),
            # value at the trial point (:292) -- one CUDA graph and one device-to-host copy; re-trials with a
            # larger M (:298-302) replay only the second half
            x_new = eng.new_vector()
            value_new, decrease, r, its, status, m_eff, gnorm, hvps, dx = eng.krylov_trial(
                first, self.subspace_dim, self.reorth, M, self.r0, self.solver_eps, 100, x_out=x_new)
            if first:
                loss.calls["gradient"] += 1
                loss.calls["hess_vec_prod"] += self.subspace_dim
                first = False
            loss.calls["value"] += 1
            loss.calls["forward"] += 1
            if status == KCRN_E_NOTPD:
                raise np.linalg.LinAlgError("T + r I is not positive definite in the cubic sub-problem")
            loss.adopt_iterate(x_new)
            loss._note_value(value_new, x_new, owned=True)    # x_new is never written again
            if not (value_new > self.value - decrease and tries < 20):   # cubic.py:297
                break
            M = M / self.beta
            tries += 1
        self.x, self.reg_coef, self.value, self.r0 = x_new, M, value_new, r
        self.solver_it += its
        self._dx = dx
        self.history.append(dict(f=value_new, M=M, solver_it=self.solver_it, m_eff=m_eff, backtracks=tries, r=r,
                                 hvps=hvps, gnorm=gnorm, dx=dx))

    def _step_on_device(self, loss, eng):
        """cubic.py:265-309 as one graph launch: gradient, Lanczos, then the trials of the line search (sub-problem,
        x + V s, value, the test of :297, M / beta) in a CUDA-graph WHILE node, the commit of (x, value, M, r0) and the
        best-point tracking of Oracle.value (loss.py:70-72) -- the host reads one record."""
        rec, status, x_new = eng.krylov_step(self.subspace_dim, self.reorth, self.value, self.reg_coef, self.beta,
                                             self.r0, self.solver_eps, 100, 20, loss.f_opt)
        if rec is None:
            loss.forget_iterate()
            if status == KCRN_E_NOTPD:
                raise np.linalg.LinAlgError("T + r I is not positive definite in the cubic sub-problem")
            raise RuntimeError(f"kcrn_krylov_run stopped with status {status}")
        tries = int(rec[13])
        loss.calls["gradient"] += 1
        loss.calls["hess_vec_prod"] += self.subspace_dim
        loss.calls["value"] += 1 + tries
        loss.calls["forward"] += 1 + tries
        loss.adopt_iterate(x_new)
        if rec[11] < loss.f_opt:                        # a trial point (accepted or not) beat the best value so far
            loss.f_opt = rec[11]
            loss._defer_x_opt(eng.best_point)           # it sits in the engine; copied when somebody reads x_opt
        self.x, self.reg_coef, self.value, self.r0 = x_new, rec[12], rec[0], rec[2]
        self.solver_it += int(rec[3])
        self._dx = rec[8]
        self.history.append(dict(f=rec[0], M=rec[12], solver_it=self.solver_it, m_eff=int(rec[5]), backtracks=tries,
                                 r=rec[2], hvps=int(rec[7]), gnorm=rec[6], dx=rec[8]))

    def run(self, x0, t_max=np.inf, it_max=np.inf, ls_it_max=None):
        if not self.device_loop:
            return super().run(x0, t_max=t_max, it_max=it_max, ls_it_max=ls_it_max)
        if t_max is np.inf and it_max is np.inf:
            it_max = 100
            print(f'{self.label}: The number of iterations is set to {it_max}.')
        self.t_max, self.it_max, self.ls_it_max = t_max, it_max, it_max
        if not self.initialized:
            self.init_run(x0)
            self.initialized = True
        loss, eng = self.loss, self.loss.engine
        if self.value is None:
            self.value = loss.value(self.x)
        else:
            loss._forward(self.x)
        while not self.check_convergence():
            left = self.it_max - self.it if self.it_max is not np.inf else self.device_chunk
            chunk = int(max(1, min(left, self.device_chunk, (2 << 30) // (8 * self.dim))))
            if self.t_max is not np.inf:
                chunk = min(chunk, 8)                  # the time budget is checked between chunks
            recs, status = eng.krylov_run(chunk, self.subspace_dim, self.reorth, self.value, self.reg_coef, self.beta,
                                          self.r0, self.solver_eps, 100, self.tolerance, 20, keep_xs=True,
                                          f_opt=loss.f_opt)
            xs, x_best, x_now = eng.run_view(len(recs))
            t_chunk = self.t
            for k, rec in enumerate(recs):
                tries = int(rec[13])
                loss.calls["gradient"] += 1
                loss.calls["hess_vec_prod"] += self.subspace_dim
                loss.calls["value"] += 1 + tries
                loss.calls["forward"] += 1 + tries
                self.reg_coef, self.value, self.r0, self._dx = rec[12], rec[0], rec[2], rec[8]
                self.solver_it += int(rec[3])
                self.history.append(dict(f=rec[0], M=rec[12], solver_it=self.solver_it, m_eff=int(rec[5]), backtracks=tries,
                                         r=rec[2], hvps=int(rec[7]), gnorm=rec[6], dx=rec[8]))
                self.x = xs[k]
                self.it += 1
                self.t = t_chunk + rec[14]
                if self.should_update_trace():
                    self.update_trace()                # clones self.x
                self.max_progress = max(self.time_progress, self.iterations_progress)
            if len(recs):
                if recs[-1][15] != 0.0:
                    loss.x_opt, loss.f_opt = x_best.clone(), float(recs[-1][11])
                self.x = x_now.clone()
                loss.adopt_iterate(self.x)
                if self.tolerance > 0:
                    self.x_old_tol = self.x            # step_length() reads the device's |x_new - x|
            if status == KCRN_E_NOTPD:
                raise np.linalg.LinAlgError("T + r I is not positive definite in the cubic sub-problem")
            self.t = max(self.t, time.perf_counter() - self.t_start)
        self.initialized = False
        return self.trace

    def step_length(self):
        return self._dx if self._dx is not None else super().step_length()

    def init_run(self, x0, *args, **kwargs):
        x0 = self.loss.to_device(x0)
        super().init_run(x0, *args, **kwargs)
        self.trace.solver_its = [0]
        self.loss.reset()
        self._dx = None

    def update_trace(self):
        super().update_trace()
        self.trace.solver_its.append(self.solver_it)


# --------------------------------------------------------------------------------- config-4 baselines
class Cubic_LS(Optimizer):
    """Cubic regularized Newton with line search on the full space (reference cubic.py:115-235).

    ``cubic_solver="full"``: dense Hessian A' D A / n formed on the FP64 tensor cores (kcrn_hessian_dense) and the
    dense secular-equation solve of ``optimizer.dense``.  ``cubic_solver="CG"`` (what the reference driver uses for
    dim >= 500, cubic_newton.py:76-82): matrix-free -- every (H + lam I)^-1 application is a conjugate-gradient solve
    whose operator is the device HVP; the CG recurrence itself follows scipy.sparse.linalg.cg (x0 = 0, atol = 0,
    rtol = solver_eps, maxiter = 10 n, no preconditioner) with torch vector operations on the device."""

    def __init__(self, reg_coef=None, cubic_solver="CG", solver_it_max=100, solver_eps=1e-8, beta=0.5, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if cubic_solver not in ("full", "CG"):
            raise ValueError("cubic_solver must be 'full' or 'CG'")
        self.cubic_solver = cubic_solver
        self.solver_it = 0
        self.solver_it_max = solver_it_max
        self.solver_eps = solver_eps
        self.beta = beta
        self.r0 = 0.1
        self.value = None
        self.cg_iterations = 0
        self.reg_coef = self.loss.hessian_lipschitz if reg_coef is None else reg_coef

    # -- matrix-free solve of (H + lam I) y = b, as scipy.sparse.linalg.cg iterates it
    def _cg(self, lam, b, rtol):
        eng = self.loss.engine
        bnrm2 = float(b.norm())
        if bnrm2 == 0:
            return b.clone()
        atol = rtol * bnrm2
        x = b.new_zeros(b.shape)
        r = b.clone()
        p, rho_prev = None, None
        for it in range(10 * b.numel()):
            if float(r.norm()) < atol:
                break
            rho_cur = float(r @ r)
            if it > 0:
                p.mul_(rho_cur / rho_prev).add_(r)
            else:
                p = r.clone()
            q = eng.hvp(p).add_(p, alpha=lam)
            alpha = rho_cur / float(p @ q)
            x.add_(p, alpha=alpha)
            r.sub_(q, alpha=alpha)
            rho_prev = rho_cur
            self.cg_iterations += 1
        return x

    def cubic_solver_root_CG(self, M, it_max=100, epsilon=1e-8, r0=0.1):
        """cubic_solver_root with conjugate-gradient solves (reference cubic.py:152-182)."""
        from optimizer.dense import newton_scalar

        g = self.grad
        self.loss._forward(self.x)
        memo = {}

        def s_of(lam):
            if lam not in memo:
                memo.clear()
                memo[lam] = self._cg(lam, -g, epsilon)
            return memo[lam]

        def func(lam):
            return lam ** 2 - M ** 2 * float(s_of(lam).norm()) ** 2

        def grad(lam):
            s = s_of(lam)
            return 2 * lam + 2 * M ** 2 * float(s @ self._cg(lam, s, epsilon))

        r, its = newton_scalar(func, grad, r0, epsilon, it_max)
        s = self._cg(r, -g, epsilon)
        ns = float(s.norm())
        return s, its, r, r / 2 * ns ** 2 - M / 3 * ns ** 3 - float(g @ s) / 2

    def cubic_solver_root_full(self, M, it_max=100, epsilon=1e-8, r0=0.1):
        from optimizer.dense import cubic_solver_dense

        return cubic_solver_dense(self.grad, self.hess, M, it_max, epsilon, r0=r0)

    def step(self):
        loss = self.loss
        if self.value is None:
            self.value = loss.value(self.x)
        self.grad = loss.gradient(self.x)
        if self.cubic_solver == "full":
            self.hess = loss.hessian(self.x)
        if loss.norm(self.grad) < self.tolerance:        # cubic.py:201-202
            return
        solve = self.cubic_solver_root_full if self.cubic_solver == "full" else self.cubic_solver_root_CG
        M = self.reg_coef * self.beta
        while True:
            s, its, r, decrease = solve(M, self.solver_it_max, self.solver_eps, r0=self.r0)
            x_new = self.x + s
            value_new = loss.value(x_new)
            if not value_new > self.value - decrease:    # unbounded backtracking, cubic.py:214
                break
            loss._forward(self.x)                        # the next solve needs the Hessian at x again
            M = M / self.beta
        self.x, self.reg_coef, self.value, self.r0 = x_new, M, value_new, r
        self.solver_it += its

    def init_run(self, x0, *args, **kwargs):
        super().init_run(self.loss.to_device(x0), *args, **kwargs)
        self.trace.solver_its = [0]
        self.loss.reset()

    def update_trace(self):
        super().update_trace()
        self.trace.solver_its.append(self.solver_it)


class SSCN(Optimizer):
    """Stochastic subspace cubic Newton, coordinate version (reference cubic.py:321-408): a random coordinate
    block I per step (same numpy Generator stream as the reference), block gradient / Hessian from the device,
    dense m x m cubic solve, incremental update of the cached A x."""

    def __init__(self, reg_coef=None, subspace_dim=100, solver_eps=1e-8, beta=0.5, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.solver_it = 0
        self.subspace_dim = subspace_dim
        self.solver_eps = solver_eps
        self.beta = beta
        self.r0 = 0.1
        self.value = None
        self.tolerance = 0                                # cubic.py:345
        self.reg_coef = self.loss.hessian_lipschitz if reg_coef is None else reg_coef

    def step(self):
        from optimizer.dense import cubic_solver_dense

        loss = self.loss
        if self.value is None:
            self.value = loss.value(self.x)
        I = self.rng.choice(self.dim, size=self.subspace_dim, replace=False)
        I_dev = loss._index_tensor(I)
        self.grad = loss.partial_gradient(self.x, I)
        self.hess = loss.partial_hessian(self.x, I)
        M = max(self.reg_coef * self.beta, np.finfo(float).eps)
        x_new = self.x.clone()
        Ax = loss._mat_vec_prod.clone()
        eps = np.finfo(float).eps
        while True:
            s, its, r, decrease = cubic_solver_dense(self.grad, self.hess, M, r0=self.r0, epsilon=eps)
            x_new[I_dev] = self.x[I_dev] + s
            loss.update_mat_vec_product(Ax, s, I)
            value_new = loss.value(x_new)
            if not value_new > self.value - decrease:
                break
            M = M / self.beta
        self.x, self.reg_coef, self.value, self.r0 = x_new, M, value_new, r
        self.solver_it += its

    def init_run(self, x0, *args, **kwargs):
        super().init_run(self.loss.to_device(x0), *args, **kwargs)
        self.trace.solver_its = [0]
        self.loss.reset()

    def update_trace(self):
        super().update_trace()
        self.trace.solver_its.append(self.solver_it)
