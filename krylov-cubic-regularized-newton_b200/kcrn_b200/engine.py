"""Thin host wrapper around one libkcrn handle (one row shard of the data matrix on one GPU).

torch is used here for device memory, streams and (in ``kcrn_b200.dist``) process-group plumbing
only: every number is produced by the CUDA kernels behind the C ABI (include/kcrn.h).
"""
from __future__ import annotations

import atexit
import ctypes
import weakref
from ctypes import byref, c_double, c_int, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import check

MAX_SUBSPACE = 512


_live = weakref.WeakSet()


@atexit.register
def _close_all():
    """Release every handle (graphs, NCCL communicator, device memory) while the interpreter is still whole."""
    for eng in list(_live):
        try:
            eng.close()
        except Exception:
            pass


class _DeviceView:
    """Exposes memory owned by the handle to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class Engine:
    """Device-resident logistic-regression problem: CSR shard, CSC copy, tile lists, caches."""

    def __init__(self, indptr, indices, data, labels01, shape, l2: float = 0.0, n_global: int | None = None,
                 device=None):
        import torch

        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("kcrn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.n, self.d = int(shape[0]), int(shape[1])
        on_device = int(hasattr(indptr, "data_ptr"))
        if on_device:
            keep = [indptr.contiguous().to(torch.int32), indices.contiguous().to(torch.int32),
                    data.contiguous().to(torch.float64), labels01.contiguous().to(torch.float64)]
            ptrs = [t.data_ptr() for t in keep]
            self.nnz = int(keep[1].numel())
        else:
            keep = [np.ascontiguousarray(indptr, np.int32), np.ascontiguousarray(indices, np.int32),
                    np.ascontiguousarray(data, np.float64), np.ascontiguousarray(labels01, np.float64)]
            ptrs = [a.ctypes.data for a in keep]
            self.nnz = int(keep[0][-1])
            if keep[0].shape[0] != self.n + 1 or keep[1].shape[0] != self.nnz or keep[2].shape[0] != self.nnz:
                raise ValueError("CSR arrays do not match the given shape")
        if keep[3].shape[0] != self.n:
            raise ValueError("labels must have one entry per row")
        self._stream = torch.cuda.current_stream().cuda_stream
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        h = c_void_p()
        check(self.lib.kcrn_create(byref(h), self.n, self.d, self.nnz, ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                                   float(l2), int(n_global if n_global is not None else self.n), on_device,
                                   self._stream))
        self.h = h
        self.l2 = float(l2)
        self.n_global = int(n_global if n_global is not None else self.n)
        self._views = None
        self._rec = (c_double * 12)()
        self._cnt = (c_int64 * 8)()
        self.world = 1
        self.p2p = False          # peer exchange attached and on (row-sharded device loop needs it)
        self._rec16 = (c_double * 16)()
        self._iters, self._status = ctypes.c_int(0), ctypes.c_int(0)
        _live.add(self)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            self.lib.kcrn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _use_stream(self):
        # the raw handle of torch's current stream (torch.cuda.current_stream() builds a Stream object: ~5 us per call)
        s = self._raw_stream(self._dev_index) if self._raw_stream else self.torch.cuda.current_stream().cuda_stream
        if s != self._stream:
            check(self.lib.kcrn_set_stream(self.h, s))
            self._stream = s

    def _vec(self, t, size, name):
        torch = self.torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
                and t.numel() == size):
            raise ValueError(f"{name}: need a contiguous float64 CUDA tensor with {size} elements")
        return t

    def new_vector(self, size=None):
        return self.torch.empty(self.d if size is None else size, dtype=self.torch.float64, device=self.device)

    def view(self, ptr, shape, typestr="<f8"):
        return self.torch.as_tensor(_DeviceView(ptr, shape, typestr, self), device=self.device)

    def sync(self):
        check(self.lib.kcrn_sync(self.h))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        check(self.lib.kcrn_comm_init(self.h, buf, int(rank), int(world)))
        self.world = int(world)

    def p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        check(self.lib.kcrn_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, all_handles: bytes):
        buf = ctypes.create_string_buffer(bytes(all_handles), len(all_handles))
        check(self.lib.kcrn_p2p_attach(self.h, buf))
        self.p2p = True

    def p2p_enable(self, on: bool):
        check(self.lib.kcrn_p2p_enable(self.h, int(bool(on))))
        self.p2p = bool(on)

    def nccl_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        check(self.lib.kcrn_nccl_unique_id(buf))
        return buf.raw

    # ------------------------------------------------------------------ oracle (optimizer/loss.py)
    def set_x(self, x):
        self._use_stream()
        check(self.lib.kcrn_set_x(self.h, self._vec(x, self.d, "x").data_ptr()))

    def set_x_only(self, x):
        """x moves, the cached A x does not (update_mat_vec_product overrides it until reset())."""
        self._use_stream()
        check(self.lib.kcrn_set_x_only(self.h, self._vec(x, self.d, "x").data_ptr()))

    def value_tensor(self, out=None):
        self._use_stream()
        out = self.new_vector(1) if out is None else out
        check(self.lib.kcrn_value(self.h, out.data_ptr()))
        return out

    def value(self) -> float:
        return float(self.value_tensor().item())

    def gradient(self, out=None, keep_internal=False):
        self._use_stream()
        if keep_internal:
            check(self.lib.kcrn_gradient(self.h, None))
            return None
        out = self.new_vector() if out is None else self._vec(out, self.d, "out")
        check(self.lib.kcrn_gradient(self.h, out.data_ptr()))
        return out

    def hvp(self, v, out=None):
        self._use_stream()
        out = self.new_vector() if out is None else self._vec(out, self.d, "out")
        check(self.lib.kcrn_hvp(self.h, self._vec(v, self.d, "v").data_ptr(), out.data_ptr()))
        return out

    def cached_Ax(self):
        p = c_void_p()
        check(self.lib.kcrn_cached_Ax(self.h, byref(p)))
        return self.view(p.value, (self.n,))

    def set_Ax(self, Ax):
        self._use_stream()
        check(self.lib.kcrn_set_Ax(self.h, self._vec(Ax, self.n, "Ax").data_ptr()))

    def value_grad_host(self, x_host: np.ndarray, want_grad=True):
        """numpy in, numpy out: H2D of x, forward sweep, value (+ gradient), D2H -- one call."""
        self._use_stream()
        x_host = np.ascontiguousarray(x_host, np.float64)
        val = c_double()
        grad = np.empty(self.d) if want_grad else None
        check(self.lib.kcrn_value_grad_host(self.h, x_host.ctypes.data, byref(val),
                                            grad.ctypes.data if want_grad else None))
        return val.value, grad

    def hvp_host(self, x_host, v_host: np.ndarray, out_host: np.ndarray | None = None):
        self._use_stream()
        v_host = np.ascontiguousarray(v_host, np.float64)
        out = np.empty(self.d) if out_host is None else out_host
        xp = None if x_host is None else np.ascontiguousarray(x_host, np.float64).ctypes.data
        check(self.lib.kcrn_hvp_host(self.h, xp, v_host.ctypes.data, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ Lanczos (optimizer/cubic.py:77-111)
    def lanczos(self, m: int, reorth: bool = False, g=None):
        self._use_stream()
        check(self.lib.kcrn_lanczos(self.h, _ptr(g), int(m), int(bool(reorth))))
        self._m = int(m)
        self._views = None

    def lanczos_view(self):
        """(V [m, d] -- row i is the i-th Lanczos vector, alphas [m], bet [m + 1] with bet[0] = |g|, bet[j + 1] = beta_j, bet[m] = last beta, info int32[8])."""
        if self._views is None:
            pv, pa, pb, pi = c_void_p(), c_void_p(), c_void_p(), c_void_p()
            check(self.lib.kcrn_lanczos_view(self.h, byref(pv), byref(pa), byref(pb), byref(pi)))
            m = self._m
            self._views = (self.view(pv.value, (m, self.d)), self.view(pa.value, (m,)), self.view(pb.value, (m + 1,)),
                           self.view(pi.value, (8,), "<i4"))
        return self._views

    def lanczos_begin(self, g, m: int, reorth: bool = False):
        self._use_stream()
        check(self.lib.kcrn_lanczos_begin(self.h, self._vec(g, self.d, "g").data_ptr(), int(m), int(bool(reorth))))
        self._m = int(m)
        self._views = None

    def lanczos_current(self):
        p = c_void_p()
        check(self.lib.kcrn_lanczos_current(self.h, byref(p)))
        return self.view(p.value, (self.d,))

    def lanczos_feed(self, hv) -> bool:
        self._use_stream()
        fin = c_int()
        check(self.lib.kcrn_lanczos_feed(self.h, self._vec(hv, self.d, "hv").data_ptr(), byref(fin)))
        return bool(fin.value)

    # ------------------------------------------------------------------ sub-problem / trial point
    def cubic(self, M: float, r0: float, eps: float = 1e-8, it_max: int = 100):
        self._use_stream()
        check(self.lib.kcrn_cubic_from_lanczos(self.h, float(M), float(r0), float(eps), int(it_max)))

    def trial_point(self, x, out=None):
        self._use_stream()
        out = self.new_vector() if out is None else self._vec(out, self.d, "out")
        check(self.lib.kcrn_trial_point(self.h, self._vec(x, self.d, "x").data_ptr(), out.data_ptr()))
        return out

    def krylov_trial(self, first: bool, m: int, reorth: bool, M: float, r0: float, eps: float = 1e-8,
                     it_max: int = 100, x_out=None):
        """One trial of a Krylov-CRN step as one CUDA graph + one D2H (kcrn_krylov_trial).  Returns the step record
        tuple of ``step_record``; ``x_out`` receives the trial point."""
        self._use_stream()
        check(self.lib.kcrn_krylov_trial(self.h, int(bool(first)), int(m), int(bool(reorth)), float(M), float(r0),
                                         float(eps), int(it_max), _ptr(x_out), self._rec))
        self._m = int(m)
        self._views = None
        r = self._rec
        return r[0], r[1], r[2], int(r[3]), int(r[4]), int(r[5]), r[6], int(r[7]), r[8]

    def krylov_run(self, max_iters: int, m: int, reorth: bool, value0: float, reg_coef: float, beta: float = 0.5,
                   r0: float = 0.1, eps: float = 1e-8, solver_it_max: int = 100, tol: float = 0.0, max_tries: int = 20,
                   keep_xs: bool = False, f_opt: float = float("inf")):
        """Up to ``max_iters`` outer iterations of Krylov CRN with the line search on the device: one graph launch
        (kcrn_krylov_run).  Returns (records [iters, 16], status)."""
        self._use_stream()
        recs = np.empty((int(max_iters), 16), dtype=np.float64)
        iters, status = ctypes.c_int(0), ctypes.c_int(0)
        check(self.lib.kcrn_krylov_run(self.h, int(max_iters), int(m), int(bool(reorth)), float(value0), float(reg_coef),
                                       float(beta), float(r0), float(eps), int(solver_it_max), float(tol), int(max_tries),
                                       int(bool(keep_xs)), float(f_opt), byref(iters), byref(status),
                                       recs.ctypes.data_as(c_void_p)))
        self._m = int(m)
        self._views = None
        return recs[:iters.value], status.value

    def krylov_step(self, m: int, reorth: bool, value0: float, reg_coef: float, beta: float, r0: float, eps: float,
                    solver_it_max: int, max_tries: int, f_opt: float):
        """ONE outer iteration with its line search on the device (kcrn_krylov_run with max_iters = 1: one graph
        launch, one synchronisation) and the new iterate copied into a fresh vector (kcrn_current_x, not waited
        for).  Returns (16-double record or None when the sub-problem failed, status, x_new)."""
        self._use_stream()
        check(self.lib.kcrn_krylov_run(self.h, 1, int(m), int(bool(reorth)), float(value0), float(reg_coef), float(beta),
                                       float(r0), float(eps), int(solver_it_max), 0.0, int(max_tries), 0, float(f_opt),
                                       byref(self._iters), byref(self._status), self._rec16))
        self._m = int(m)
        self._views = None
        x_new = self.new_vector()
        check(self.lib.kcrn_current_x(self.h, x_new.data_ptr()))
        return (list(self._rec16) if self._iters.value == 1 else None), self._status.value, x_new

    def best_point(self):
        """Clone of the best point the device loop has tracked (kcrn_run_view)."""
        return self.run_view(0)[1].clone()

    def run_view(self, iters: int):
        """(accepted iterates [iters, d] or None, best point [d], current iterate [d]) -- device views."""
        pxs, pb, px = c_void_p(), c_void_p(), c_void_p()
        check(self.lib.kcrn_run_view(self.h, byref(pxs), byref(pb), byref(px)))
        xs = self.view(pxs.value, (int(iters), self.d)) if pxs.value and iters > 0 else None
        return xs, self.view(pb.value, (self.d,)), self.view(px.value, (self.d,))

    def values_batch(self, X):
        """f(x_j) for the rows of X [k, d] (device tensor) -- one sweep of A per 8 iterates (kcrn_values_batch)."""
        self._use_stream()
        torch = self.torch
        X = X.contiguous()
        if X.dim() != 2 or X.shape[1] != self.d or X.dtype != torch.float64 or not X.is_cuda:
            raise ValueError(f"values_batch expects a float64 CUDA tensor [k, {self.d}]")
        out = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
        check(self.lib.kcrn_values_batch(self.h, X.data_ptr(), int(X.shape[0]), out.data_ptr()))
        return out

    def step_record(self, with_value: bool = True):
        """(value, model_decrease, r, newton_its, status, m_eff, |g|, hvps, |x_trial - x|) -- one D2H, synchronous."""
        self._use_stream()
        check(self.lib.kcrn_step_record(self.h, int(with_value), self._rec))
        r = self._rec
        return r[0], r[1], r[2], int(r[3]), int(r[4]), int(r[5]), r[6], int(r[7]), r[8]

    def axpy_basis(self, x, V, s, out=None):
        """out = x + sum_i s[i] V[i]  with V given as [m, d] rows."""
        self._use_stream()
        out = self.new_vector() if out is None else out
        m = int(s.numel())
        check(self.lib.kcrn_axpy_basis(x.data_ptr(), V.data_ptr(), s.data_ptr(), m, self.d, out.data_ptr(),
                                       self._stream))
        return out

    # ------------------------------------------------------------------ config-4 rows
    def hessian_dense(self, out=None):
        self._use_stream()
        out = self.torch.empty((self.d, self.d), dtype=self.torch.float64, device=self.device) if out is None else out
        check(self.lib.kcrn_hessian_dense(self.h, out.data_ptr()))
        return out

    def partial_gradient(self, I):
        self._use_stream()
        out = self.new_vector(int(I.numel()))
        check(self.lib.kcrn_partial_gradient(self.h, I.data_ptr(), int(I.numel()), out.data_ptr()))
        return out

    def partial_hessian(self, I):
        self._use_stream()
        m = int(I.numel())
        out = self.torch.empty((m, m), dtype=self.torch.float64, device=self.device)
        check(self.lib.kcrn_partial_hessian(self.h, I.data_ptr(), m, out.data_ptr()))
        return out

    def update_Ax(self, I, delta):
        self._use_stream()
        check(self.lib.kcrn_update_Ax(self.h, I.data_ptr(), delta.data_ptr(), int(I.numel())))

    # ------------------------------------------------------------------ inspection
    def csc(self):
        indptr = np.empty(self.d + 1, np.int32)
        indices = np.empty(self.nnz, np.int32)
        data = np.empty(self.nnz, np.float64)
        check(self.lib.kcrn_csc_view(self.h, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data))
        return indptr, indices, data

    def counters(self, reset: bool = False):
        check(self.lib.kcrn_counters(self.h, self._cnt, int(reset)))
        keys = ("kernels", "graph_launches", "hvps", "row_sweeps", "col_sweeps", "allreduces", "h2d_bytes", "d2h_bytes")
        return dict(zip(keys, (int(v) for v in self._cnt)))

    def time_sweep(self, which: int, reps: int) -> float:
        self._use_stream()
        us = c_double()
        check(self.lib.kcrn_time_sweep(self.h, int(which), int(reps), byref(us)))
        return us.value

    def flush_l2(self, nbytes: int = 256 << 20) -> None:
        """Benchmark utility: everything out of L2 (allocate, write, free ``nbytes`` > L2; returns when done)."""
        self._use_stream()
        check(self.lib.kcrn_flush_l2(self.h, ctypes.c_size_t(int(nbytes))))

    def gather_variants(self) -> dict:
        """Gather variant of the sweep kernel measured fastest per matrix and class of pass (0 stored order,
        1 gather-ordered tiles, 2 gather-ordered tiles + LDGSTS gathers); same bits whichever runs."""
        out = (c_int * 4)()
        check(self.lib.kcrn_gather_variants(self.h, out))
        return {"csr_plain": out[0], "csr_lanczos": out[1], "csc_plain": out[2], "csc_lanczos": out[3]}

    # algorithmic bytes of one pass; used by bench.py's roofline
    def sweep_bytes(self, which: int) -> int:
        if which == 0:   # CSR: 12 nnz + 4(n+1) + 8d (gather v) + 8n (weights) + 8n (write t)
            return 12 * self.nnz + 4 * (self.n + 1) + 8 * self.d + 16 * self.n
        return 12 * self.nnz + 4 * (self.d + 1) + 8 * self.n + 16 * self.d  # CSC: gather t, read v, write out

    def hvp_bytes(self) -> int:
        return self.sweep_bytes(0) + self.sweep_bytes(1)


def cubic_tridiag(alphas, betas, gnorm, M, r0=0.1, eps=1e-8, it_max=100, device=None):
    """Stand-alone sub-problem solve (kcrn_cubic_tridiag): T = tridiag(betas, alphas, betas), g = gnorm e1.

    Returns (s tensor [m], r, model_decrease, newton_iterations, status)."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("kcrn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def put(a):
        if hasattr(a, "data_ptr"):
            return a.to(device=dev, dtype=torch.float64).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a, np.float64), device=dev)

    a = put(alphas)
    m = int(a.numel())
    if m < 1 or m > MAX_SUBSPACE:
        raise ValueError(f"sub-problem dimension {m} outside [1, {MAX_SUBSPACE}]")
    b = put(betas) if m > 1 else torch.zeros(1, dtype=torch.float64, device=dev)
    if m > 1 and int(b.numel()) != m - 1:
        raise ValueError("betas must have m - 1 entries")
    out = torch.empty(m + 4, dtype=torch.float64, device=dev)
    check(lib.kcrn_cubic_tridiag(a.data_ptr(), b.data_ptr(), m, float(gnorm), float(M), float(r0), float(eps),
                                 int(it_max), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    tail = out[m:].cpu().numpy()
    return out[:m], float(tail[0]), float(tail[1]), int(tail[2]), int(tail[3])


DENSE_KERNEL_MAX = 64


def cubic_dense_small(H, g, M, r0=0.1, eps=1e-8, it_max=100):
    """Dense m x m sub-problem, m <= 64, in one kernel (kcrn_cubic_dense).  H [m, m] and g [m] are float64 CUDA tensors.
    Returns (s tensor [m], r, model_decrease, newton_iterations, status) after ONE device-to-host copy."""
    import torch

    lib = _lib.load()
    m = int(g.numel())
    H = H.contiguous()
    g = g.contiguous()
    out = torch.empty(m + 4, dtype=torch.float64, device=g.device)
    check(lib.kcrn_cubic_dense(H.data_ptr(), g.data_ptr(), m, float(M), float(r0), float(eps), int(it_max), out.data_ptr(),
                               torch.cuda.current_stream().cuda_stream))
    tail = out[m:].cpu().numpy()
    return out[:m], float(tail[0]), float(tail[1]), int(tail[2]), int(tail[3])
