"""Seeded synthetic sparse logistic-regression problems.

Real LIBSVM files (w8a, rcv1_train.binary, news20.binary; reference
``cubic_newton.py:43-52``) cannot be downloaded here, so every test, golden fixture and
bench line uses matrices of the same *shape* produced by this generator:

* row lengths  ``clip(Poisson(nnz/n), 1, d)``
* column popularity ``p_j ~ rank_j**(-a)`` under a random permutation (power-law columns)
* distinct, sorted int32 column indices inside every row
* values: all ones (w8a is binary) or ``lognormal(0, 0.5)`` followed by unit-L2 row scaling
  (rcv1/news20 are row-normalised tf-idf)
* labels ``sign(A x* + 0.3 eps - median)`` in {-1,+1}

The generator is pure numpy (PCG64 ``default_rng``) so the container that writes the golden
fixtures and the GPU box that replays them see bit-identical inputs; ``fingerprint`` lets the
tests assert that.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (n, d, nnz target, values, default subspace dims)
    "tiny": (257, 61, 2_000, "lognormal", (4,)),
    "small": (3_000, 500, 60_000, "lognormal", (10,)),
    "w8a": (49_749, 300, 1_200_000, "ones", (10,)),
    "rcv1": (20_242, 47_236, 1_500_000, "lognormal", (10, 50)),
    "news20": (19_996, 1_355_191, 9_000_000, "lognormal", (50,)),
}


@dataclass
class Problem:
    name: str
    n: int
    d: int
    indptr: np.ndarray   # int32 [n+1]
    indices: np.ndarray  # int32 [nnz]
    data: np.ndarray     # float64 [nnz]
    b: np.ndarray        # float64 [n] in {-1,+1}
    x0: np.ndarray       # float64 [d]  (0.5 * ones, cubic_newton.py:61)

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])

    def scipy_csr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=(self.n, self.d))

    def row_slice(self, lo: int, hi: int) -> "Problem":
        """Contiguous row block [lo, hi) as an independent problem (multi-GPU shards)."""
        p0, p1 = int(self.indptr[lo]), int(self.indptr[hi])
        return Problem(self.name, hi - lo, self.d,
                       (self.indptr[lo:hi + 1] - p0).astype(np.int32),
                       self.indices[p0:p1].copy(), self.data[p0:p1].copy(),
                       self.b[lo:hi].copy(), self.x0)

    def fingerprint(self) -> str:
        h = hashlib.sha256()
        for a in (self.indptr, self.indices, self.data, self.b):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def _distinct_columns(rng, cdf, want, d):
    """``want`` distinct columns drawn from the categorical with CDF ``cdf`` (sorted)."""
    got = np.empty(0, dtype=np.int64)
    draw = want + (want >> 1) + 8
    while got.size < want:
        u = rng.random(draw)
        cand = np.searchsorted(cdf, u, side="right")
        np.minimum(cand, d - 1, out=cand)
        got = np.union1d(got, cand)
        draw *= 2
    if got.size > want:
        keep = rng.permutation(got.size)[:want]
        got = np.sort(got[keep])
    return got


def make_problem(name="rcv1", n=None, d=None, nnz=None, values=None, zipf_a=1.0, seed=0) -> Problem:
    """Build one synthetic problem; ``name`` picks defaults from ``SHAPES``."""
    if name in SHAPES:
        n0, d0, nnz0, val0, _ = SHAPES[name]
    else:
        n0 = d0 = nnz0 = val0 = None
    n = int(n if n is not None else n0)
    d = int(d if d is not None else d0)
    nnz = int(nnz if nnz is not None else nnz0)
    values = values or val0 or "lognormal"

    rng = np.random.default_rng(seed)
    lens = np.clip(rng.poisson(nnz / n, size=n), 1, d).astype(np.int64)
    pop = np.arange(1, d + 1, dtype=np.float64) ** (-float(zipf_a))
    pop = pop[rng.permutation(d)]
    cdf = np.cumsum(pop / pop.sum())
    cdf[-1] = 1.0

    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    for i in range(n):
        indices[indptr[i]:indptr[i + 1]] = _distinct_columns(rng, cdf, int(lens[i]), d)

    if values == "ones":
        data = np.ones(indices.size, dtype=np.float64)
    else:
        data = rng.lognormal(0.0, 0.5, size=indices.size)
        sq = np.add.reduceat(data * data, indptr[:-1])
        data /= np.repeat(np.sqrt(sq), lens)

    x_star = rng.standard_normal(d)
    import scipy.sparse as sp
    A = sp.csr_matrix((data, indices, indptr.astype(np.int32)), shape=(n, d))
    score = A @ x_star + 0.3 * rng.standard_normal(n)
    score -= np.median(score)
    b = np.where(score > 0, 1.0, -1.0)
    return Problem(name, n, d, indptr.astype(np.int32), indices, data, b, 0.5 * np.ones(d))


def balanced_row_cuts(indptr: np.ndarray, parts: int) -> np.ndarray:
    """Row boundaries of ``parts`` contiguous blocks holding ~equal nnz."""
    nnz = int(indptr[-1])
    n = indptr.size - 1
    targets = (np.arange(1, parts, dtype=np.float64) * nnz / parts)
    inner = np.searchsorted(indptr, targets, side="left")
    cuts = np.concatenate(([0], inner, [n])).astype(np.int64)
    return np.maximum.accumulate(np.minimum(cuts, n))


def make_problem_device(n: int, d: int, nnz_per_row: float, zipf_a: float = 1.0, seed: int = 0, device="cuda",
                        perm_seed: int | None = None):
    """Large synthetic problem generated ON the GPU with torch (BASELINE config 5 scale: host generation of 1e9
    non-zeros is impractical).  Same recipe as ``make_problem`` -- Poisson row lengths, power-law column popularity
    under a random permutation, distinct sorted columns per row, lognormal values with unit-L2 rows -- but not
    bit-identical to it.  ``perm_seed``: seed of the column-popularity permutation when several shards of one matrix are
    generated with different ``seed``s.  Returns (indptr int32 [n+1], indices int32 [nnz], data float64 [nnz], labels01 float64 [n])
    as CUDA tensors."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    lens = torch.poisson(torch.full((n,), float(nnz_per_row), device=device), generator=gen).clamp_(1, d).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n, device=device, dtype=torch.int64), lens)
    u = torch.rand(rows.numel(), device=device, generator=gen, dtype=torch.float64)
    if zipf_a == 1.0:
        rank = torch.exp(u * float(np.log(d + 1.0))) - 1.0            # P(rank <= r) ~ log(r + 1) / log(d + 1)
    else:
        e = 1.0 - zipf_a
        rank = ((u * ((d + 1.0) ** e - 1.0) + 1.0) ** (1.0 / e)) - 1.0
    del u
    rank = rank.to(torch.int64).clamp_(0, d - 1)
    if perm_seed is None:
        perm = torch.randperm(d, device=device, generator=gen)
    else:
        # row shards of ONE matrix (every rank generates its own rows): the popularity of a column is a property of the
        # whole matrix, so all shards must draw it from the same permutation
        pg = torch.Generator(device=device)
        pg.manual_seed(perm_seed)
        perm = torch.randperm(d, device=device, generator=pg)
    key = rows * d + perm[rank]
    del rows, rank
    key = torch.unique(key)                                           # sorted, duplicates inside a row removed
    rows = torch.div(key, d, rounding_mode="floor")
    indices = (key - rows * d).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, device=device, dtype=torch.int64)
    torch.cumsum(counts, 0, out=indptr[1:])
    data = torch.exp(0.5 * torch.randn(indices.numel(), device=device, generator=gen, dtype=torch.float64))
    sq = torch.zeros(n, device=device, dtype=torch.float64).index_add_(0, rows, data * data)
    data /= torch.sqrt(sq)[rows]
    labels = (torch.rand(n, device=device, generator=gen) > 0.5).to(torch.float64)
    return indptr.to(torch.int32), indices, data, labels
