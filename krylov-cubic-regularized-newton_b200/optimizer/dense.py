"""Dense cubic sub-problem for the config-4 baselines (SSCN blocks, full-Hessian CRN).

Same Newton iteration on the secular equation as ``cubic_solver_root`` (reference optimizer/cubic.py:40-75,
scipy's scalar ``newton`` loop, scipy/optimize/_zeros_py.py:319-358), with the dense SPD solves
(``scipy.linalg.solve(assume_a='pos')`` = LAPACK dposv in the reference) done by cuSOLVER's Cholesky through
torch.linalg on the device.  This is library code on a baseline-only path: the hot path's sub-problem is the
tridiagonal single-CTA kernel (csrc/kcrn_cubic.cu); what config 4 accelerates by hand is the formation of
A' D A (csrc/kcrn_config4.cu).
"""
from __future__ import annotations

import numpy as np

from kcrn_b200.engine import DENSE_KERNEL_MAX, cubic_dense_small


def newton_scalar(func, fprime, x0, tol, maxiter):
    """Newton-Raphson as scipy's scalar ``newton`` iterates it for root_scalar(method='newton', xtol=tol)
    (scipy/optimize/_zeros_py.py:319-358, disp=False, rtol=0).  Returns (root, iterations)."""
    p0 = p = float(x0)
    for itr in range(maxiter):
        fval = func(p0)
        if fval == 0:
            return p0, itr
        fder = fprime(p0)
        if fder == 0:
            return p0, itr + 1
        p = p0 - fval / fder
        if abs(p - p0) <= tol:
            return p, itr + 1
        p0 = p
    return p, maxiter


def cubic_solver_dense(g, H, M, it_max=100, epsilon=1e-8, r0=0.1):
    """argmin_s <g, s> + 1/2 <s, H s> + M/3 |s|^3 for a dense symmetric H (CUDA tensors or numpy).
    Returns (s, newton_iterations, r, model_decrease); s is a CUDA tensor if g was one."""
    import torch

    as_numpy = not hasattr(g, "data_ptr")
    dev = torch.device("cuda", torch.cuda.current_device())
    gt = torch.as_tensor(np.ascontiguousarray(g, np.float64), device=dev) if as_numpy else g.to(torch.float64)
    Ht = torch.as_tensor(np.ascontiguousarray(H, np.float64), device=dev) if not hasattr(H, "data_ptr") else H
    m = gt.numel()
    Ht = Ht.reshape(m, m)
    if m <= DENSE_KERNEL_MAX:
        # small blocks (SSCN m = 10, 50): the whole Newton loop in one kernel, one device-to-host copy
        s, r, dec, its, status = cubic_dense_small(Ht, gt, float(M), r0=float(r0), eps=float(epsilon), it_max=int(it_max))
        if status != 0:
            raise np.linalg.LinAlgError("H + r I is not positive definite")
        return (s.cpu().numpy() if as_numpy else s), its, r, dec
    diag0 = Ht.diagonal().clone()
    work = Ht if not as_numpy and Ht.numel() > (1 << 26) else Ht.clone()   # huge H: shift its diagonal in place

    def factor(lam):
        work.diagonal().copy_(diag0 + lam)
        L, info = torch.linalg.cholesky_ex(work)
        if int(info.item()) != 0:
            if work is Ht:
                work.diagonal().copy_(diag0)
            raise np.linalg.LinAlgError("H + r I is not positive definite")
        return L

    def solve(L, rhs):
        return torch.cholesky_solve(rhs[:, None], L)[:, 0]

    M = float(M)
    p0 = p = float(r0)
    its, done = it_max, False
    for itr in range(it_max):
        # one host round trip per evaluation: factorisation status, |s|^2 and s.(H + p0 I)^-1 s travel together
        work.diagonal().copy_(diag0 + p0)
        L, info = torch.linalg.cholesky_ex(work)
        s = -solve(L, gt)
        q = solve(L, s)
        bad, ns2, sdot = torch.stack([info.to(torch.float64).reshape(()), s @ s, s @ q]).tolist()
        if bad != 0:
            if work is Ht:
                work.diagonal().copy_(diag0)
            raise np.linalg.LinAlgError("H + r I is not positive definite")
        fval = p0 ** 2 - M ** 2 * float(np.sqrt(ns2)) ** 2
        if fval == 0:
            p, its, done = p0, itr, True
            break
        fder = 2 * p0 + 2 * M ** 2 * sdot
        if fder == 0:
            p, its, done = p0, itr + 1, True
            break
        p = p0 - fval / fder
        if abs(p - p0) <= epsilon:
            its, done = itr + 1, True
            break
        p0 = p
    r = p
    s = -solve(factor(r), gt)
    if work is Ht:
        work.diagonal().copy_(diag0)
    ns = float(s.norm().item())
    dec = r / 2 * ns ** 2 - M / 3 * ns ** 3 - float((gt @ s).item()) / 2
    return (s.cpu().numpy() if as_numpy else s), its, r, dec
