"""LIBSVM / svmlight files -> scipy CSR, with the return contract of sklearn.datasets.load_svmlight_file (what the
reference driver calls, cubic_newton.py:52): ``(X, y)`` with X a float64 CSR matrix with sorted indices and y the
float64 labels; ``.gz`` / ``.bz2`` are decompressed by extension; ``zero_based="auto"`` treats the file as
one-based unless a feature 0 occurs.  The text is parsed by libkcrn's host-side parser (all host threads); an
optional ``.kcrn.npz`` cache beside the file skips the parse next time.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib


def _read_bytes(path: str) -> bytes:
    ext = os.path.splitext(path)[1]
    if ext == ".gz":
        import gzip
        with gzip.open(path, "rb") as f:
            return f.read()
    if ext == ".bz2":
        import bz2
        with bz2.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def parse_svmlight(text: bytes, threads: int = 0):
    """(labels f64[n], indptr i64[n+1], indices i32[nnz] as written, data f64[nnz], min_index, max_index)."""
    from ctypes import byref, c_int32, c_int64

    lib = _lib.load()
    n, nnz, lo, hi = c_int64(0), c_int64(0), c_int32(0), c_int32(-1)
    _lib.check(lib.kcrn_svmlight_count(text, len(text), int(threads), byref(n), byref(nnz), byref(lo), byref(hi)))
    labels = np.empty(n.value, np.float64)
    indptr = np.zeros(n.value + 1, np.int64)
    indices = np.empty(nnz.value, np.int32)
    data = np.empty(nnz.value, np.float64)
    _lib.check(lib.kcrn_svmlight_parse(text, len(text), int(threads), n.value, nnz.value, labels.ctypes.data,
                                       indptr.ctypes.data, indices.ctypes.data, data.ctypes.data))
    return labels, indptr, indices, data, lo.value, hi.value


def load_svmlight_file(path, n_features=None, zero_based="auto", cache=False, threads: int = 0):
    import scipy.sparse as sp

    path = os.fspath(path)
    cache_path = path + ".kcrn.npz"
    if cache and os.path.exists(cache_path) and os.path.getmtime(cache_path) >= os.path.getmtime(path):
        z = np.load(cache_path)
        labels, indptr, indices, data, lo, hi = (z["labels"], z["indptr"], z["indices"], z["data"], int(z["lo"]),
                                                 int(z["hi"]))
    else:
        labels, indptr, indices, data, lo, hi = parse_svmlight(_read_bytes(path), threads)
        if cache:
            np.savez(cache_path, labels=labels, indptr=indptr, indices=indices, data=data, lo=lo, hi=hi)
    if zero_based is False or (zero_based == "auto" and len(indices) and lo > 0):
        indices = indices - 1
        hi -= 1
        if len(indices) and lo - 1 < 0:
            raise ValueError("feature index 0 in a file declared one-based")
    n_f = (hi if len(indices) else 0) + 1
    if n_features is None:
        n_features = n_f
    elif n_features < n_f:
        raise ValueError(f"n_features was set to {n_features}, but input file contains {n_f} features")
    X = sp.csr_matrix((data, indices, indptr), shape=(len(labels), n_features))
    X.sort_indices()
    return X, labels
