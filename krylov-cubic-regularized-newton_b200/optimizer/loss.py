"""Loss oracle with the reference's surface (optimizer/loss.py in the reference), backed by libkcrn.

``LogisticRegression(A, b, store_mat_vec_prod=..., l1=0, l2=0)`` keeps the constructor, attributes
and methods every optimizer of the reference touches:

    value  gradient  hess_vec_prod  hessian  partial_gradient  partial_hessian
    mat_vec_product  update_mat_vec_product  reset  norm  inner_prod  is_equal
    n  dim  f_opt  x_opt  regularizer  l1  l2  b

but the data matrix lives on one B200 as a CSR shard plus a device-built CSC copy and all arithmetic
runs in the CUDA kernels of libkcrn.so.  Iterates may be ``torch.float64`` CUDA tensors (results stay
on the device) or numpy arrays (results come back as numpy, one H2D/D2H per call).

Reference lines mirrored: labels loss.py:190-207, value :66-73 / :215-220, gradient :223-232,
A x memo :266-286, Hessian-vector product :289-302, block methods :234-264, :279-281.
"""
from __future__ import annotations

import copy
import warnings

import numpy as np

from kcrn_b200.engine import Engine


def _is_tensor(x) -> bool:
    return hasattr(x, "data_ptr")


def label_encoding(b):
    """The {0, 1} encoding the reference picks for a label vector (loss.py:190-207), as a plain tuple -- compute it ONCE
    on the global labels of a row-sharded problem and hand it to every shard's LogisticRegression(label_encoding=...):
    a shard that holds a single class, or whose first label differs from the global first label, would otherwise encode
    differently from its peers."""
    b = np.asarray(b)
    classes = np.unique(b)
    if classes.size > 2:
        raise ValueError("The number of classes must be no more than 2 for binary classification.")
    return (tuple(float(c) for c in classes), float(b[0]) if b.size else 0.0)


def map_labels(b, encoding=None):
    """Binary labels -> {0, 1} the way the reference does it (loss.py:190-207); `encoding` = label_encoding(global b)."""
    b = np.asarray(b)
    if encoding is None:
        classes, first = np.unique(b), (b[0] if b.size else 0.0)
    else:
        classes, first = np.asarray(encoding[0]), encoding[1]
        if np.setdiff1d(np.unique(b), classes).size:
            raise ValueError("labels outside the given encoding")
    if classes.size == 1:
        warnings.warn("The labels have only one unique value.")
    if classes.size > 2:
        raise ValueError("The number of classes must be no more than 2 for binary classification.")
    if classes.size == 2 and not np.array_equal(classes, [0, 1]):
        if np.array_equal(classes, [1, 2]):
            return (b - 1).astype(np.float64)
        if np.array_equal(classes, [-1, 1]):
            return ((b + 1) / 2).astype(np.float64)
        return 1.0 * (b == first)
    return b.astype(np.float64)


class Oracle:
    """Base of all objectives: regularisation bookkeeping and best-value tracking (loss.py:29-73)."""

    def __init__(self, l1=0, l2=0, l2_in_prox=False, regularizer=None, seed=42):
        if l1 < 0.0:
            raise ValueError("Invalid value for l1 regularization: {}".format(l1))
        if l2 < 0.0:
            raise ValueError("Invalid value for l2 regularization: {}".format(l2))
        if l1 > 0 or l2_in_prox or regularizer is not None:
            raise NotImplementedError("prox-regularised objectives are outside the accelerated path "
                                      "(the reference driver always uses l1 = 0, l2_in_prox = False)")
        self.l1, self.l2, self.l2_in_prox = l1, l2, False
        self.regularizer = None
        self.x_opt, self.f_opt = None, np.inf
        self.seed = seed
        self.rng = np.random.default_rng(seed)

    def set_seed(self, seed):
        self.seed, self.rng = seed, np.random.default_rng(seed)

    # x_opt (loss.py:70-72) is a plain attribute for every writer; a fused device step that tracked the best point in
    # the engine's own buffer leaves a callable instead, and the copy is made when somebody reads x_opt
    @property
    def x_opt(self):
        if self._x_opt_source is not None:
            self._x_opt, self._x_opt_source = self._x_opt_source(), None
        return self._x_opt

    @x_opt.setter
    def x_opt(self, x):
        self._x_opt, self._x_opt_source = x, None

    def _defer_x_opt(self, source):
        self._x_opt, self._x_opt_source = None, source

    def _note_value(self, value, x, owned=False):
        """loss.py:70-72.  owned: x is a tensor the caller created for this value and will never write again (the trial
        point of an optimizer step): the copy is then made when somebody reads x_opt, not on the step's critical path."""
        if value < self.f_opt:
            if owned and _is_tensor(x):
                version = x._version

                def copy_now():
                    if x._version != version:       # same guard as the A x memo: in-place writes bump the version
                        raise RuntimeError("x_opt: the best iterate was modified in place before x_opt was read; "
                                           "read loss.x_opt first or rebind the optimizer's x instead of writing into it")
                    return x.clone()
                self._defer_x_opt(copy_now)
            else:
                self.x_opt = x.clone() if _is_tensor(x) else copy.deepcopy(x)
            self.f_opt = value

    def value(self, x):
        value = self._value(x)
        self._note_value(value, x)
        return value


class LogisticRegression(Oracle):
    """f(x) = mean((1 - b) .* Ax - logsig(Ax)) + l2/2 |x|^2 on a sparse A, evaluated on the GPU."""

    def __init__(self, A, b, store_mat_vec_prod=False, *args, device=None, n_global=None, label_encoding=None, **kwargs):
        super().__init__(*args, **kwargs)
        import scipy.sparse as sp

        A = sp.csr_matrix(A) if not sp.issparse(A) or A.format != "csr" else A
        if A.dtype != np.float64:
            A = A.astype(np.float64)
        self.A = A
        self.b = map_labels(b, label_encoding)      # row shards: pass label_encoding(global labels)
        self.store_mat_vec_prod = store_mat_vec_prod
        self.reuse = False
        self.n, self.dim = A.shape
        self.engine = Engine(A.indptr, A.indices, A.data, self.b, A.shape, l2=self.l2, n_global=n_global,
                             device=device)
        self.torch = self.engine.torch
        self.device = self.engine.device
        self._x_key = None        # identity of the iterate the device caches belong to
        self._x_host = None       # numpy copy of it when the caller works with numpy
        self._x_ref = None        # the remembered tensor itself: its address cannot be recycled while the key lives
        self.calls = {"value": 0, "gradient": 0, "hess_vec_prod": 0, "forward": 0}

    # ------------------------------------------------------------------ iterate handling
    def to_device(self, x):
        if _is_tensor(x):
            return x if x.is_cuda else x.to(self.device)
        return self.torch.as_tensor(np.ascontiguousarray(x, np.float64), device=self.device)

    def _same_iterate(self, x) -> bool:
        if self._x_key is None:
            return False
        if _is_tensor(x):
            # same tensor object's storage, unmodified since it was remembered.  The remembered tensor is kept alive
            # (self._x_ref), so the caching allocator cannot hand its address to a new temporary while the key lives.
            return self._x_key == (x.data_ptr(), x._version, tuple(x.shape))
        return self._x_host is not None and np.array_equal(x, self._x_host)

    def _remember(self, x):
        if _is_tensor(x):
            self._x_key, self._x_host, self._x_ref = (x.data_ptr(), x._version, tuple(x.shape)), None, x
        else:
            self._x_key, self._x_host, self._x_ref = ("host",), np.array(x, dtype=np.float64, copy=True), None

    def _forward(self, x):
        """Make the device caches (A x, sigmoid, weights, residual, loss terms) belong to x (loss.py:266-277)."""
        if self.store_mat_vec_prod and self.reuse:
            # update_mat_vec_product overrides A x until reset(); the l2 terms still use the x passed in
            # (loss.py:217 value, :247 partial_gradient), so the engine's copy of x follows it
            if not self._same_iterate(x):
                self.engine.set_x_only(self.to_device(x).contiguous())
                self._remember(x)
            return
        if self.store_mat_vec_prod and self._same_iterate(x):
            return
        self.engine.set_x(self.to_device(x).contiguous())
        self.calls["forward"] += 1
        self._remember(x)

    def adopt_iterate(self, x):
        """The engine already holds the caches of tensor x (set by a fused step that wrote x through its raw
        pointer: the key is taken AFTER that write)."""
        self._remember(x)

    def forget_iterate(self):
        """The engine's caches no longer belong to any tensor the caller holds (it wrote through a raw pointer)."""
        self._x_key, self._x_host, self._x_ref = None, None, None

    def _out(self, t, like):
        return t if _is_tensor(like) else t.cpu().numpy()

    # ------------------------------------------------------------------ the reference's methods
    def mat_vec_product(self, x):
        self._forward(x)
        return self._out(self.engine.cached_Ax(), x)

    def values_of_iterates(self, xs):
        """[f(x) for x in xs] with the book-keeping of value() (calls, f_opt / x_opt; loss.py:66-73), evaluated in
        batches on the device.  The cached A x of the current iterate is left alone."""
        if len(xs) == 0:
            return np.asarray([])
        X = self.torch.stack([self.to_device(x) for x in xs])
        vals = self.engine.values_batch(X).cpu().numpy()
        for v, x in zip(vals, xs):
            self.calls["value"] += 1
            self._note_value(float(v), x)
        return vals

    def _value(self, x):
        self.calls["value"] += 1
        self._forward(x)
        return self.engine.value()

    def gradient(self, x):
        self.calls["gradient"] += 1
        self._forward(x)
        return self._out(self.engine.gradient(), x)

    def hess_vec_prod(self, x, v, grad_dif=False, eps=None):
        if grad_dif:
            g0 = self.gradient(x)
            return (self.gradient(x + eps * v) - g0) / eps
        self.calls["hess_vec_prod"] += 1
        self._forward(x)
        return self._out(self.engine.hvp(self.to_device(v).contiguous()), v)

    def hessian(self, x):
        self._forward(x)
        return self._out(self.engine.hessian_dense(), x)

    def _index_tensor(self, I):
        I = np.ascontiguousarray(np.asarray(I), np.int64)
        if I.size and (I.min() < 0 or I.max() >= self.dim):
            raise IndexError(f"coordinate block outside [0, {self.dim})")      # the kernels do not range-check
        return self.torch.as_tensor(I, device=self.device)

    def partial_gradient(self, x, I):
        self._forward(x)
        return self._out(self.engine.partial_gradient(self._index_tensor(I)), x)

    def partial_hessian(self, x, I):
        self._forward(x)
        return self._out(self.engine.partial_hessian(self._index_tensor(I)), x)

    def update_mat_vec_product(self, Ax, delta, I):
        """Cached A x <- Ax + A[:, I] delta; the cache then overrides x until reset() (loss.py:279-286)."""
        self.engine.set_Ax(self.to_device(Ax).contiguous())
        self.engine.update_Ax(self._index_tensor(I), self.to_device(delta).contiguous())
        self.reuse = True

    @property
    def _mat_vec_prod(self):
        return self.engine.cached_Ax()

    def reset(self):
        self.reuse = False
        self._x_key, self._x_host, self._x_ref = None, None, None

    # ------------------------------------------------------------------ small helpers of the surface
    @staticmethod
    def norm(x):
        if _is_tensor(x):
            return float(x.norm().item())
        return float(np.linalg.norm(x))

    @staticmethod
    def inner_prod(x, y):
        if _is_tensor(x):
            return float((x @ y).item())
        return float(np.dot(x, y))

    @staticmethod
    def is_equal(x, y):
        if x is None or y is None:
            return x is None and y is None
        if _is_tensor(x) and _is_tensor(y):
            return bool((x.shape == y.shape) and bool((x == y).all().item()))
        return np.array_equal(np.asarray(x), np.asarray(y))

    @property
    def hessian_lipschitz(self):
        raise NotImplementedError("pass reg_coef explicitly (the reference driver uses reg_coef=1e-3); the "
                                  "svds-based smoothness estimators are outside the accelerated path")
