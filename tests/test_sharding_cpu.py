"""CPU (gloo, world_size 2): the row-sharding used for multi-GPU runs -- nnz-balanced contiguous row blocks,
one allreduce of the d-vector per transpose sweep -- reproduces the single-shard numbers.  The GPU kernels
are replaced by the CPU oracle here; what is tested is the host-side partitioning and the reduction algebra
(SURVEY.md section 8(e)).  The real thing is tests/mgpu_worker.py under torchrun on >= 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kcrn_b200.dist import shard_rows
from kcrn_b200.synth import balanced_row_cuts, make_problem
from oracle import oracle as O


def test_balanced_row_cuts_cover_and_balance():
    P = make_problem("small", seed=0)
    for parts in (1, 2, 3, 8):
        cuts = balanced_row_cuts(P.indptr, parts)
        assert cuts[0] == 0 and cuts[-1] == P.n and np.all(np.diff(cuts) >= 0) and len(cuts) == parts + 1
        nnz = np.diff(P.indptr[cuts])
        assert nnz.sum() == P.nnz and nnz.max() - nnz.min() <= 2 * np.diff(P.indptr).max()
    shards = [shard_rows(P, r, 3) for r in range(3)]
    assert sum(s.n for s in shards) == P.n and sum(s.nnz for s in shards) == P.nnz
    assert np.array_equal(np.concatenate([s.indices for s in shards]), P.indices)
    assert np.array_equal(np.concatenate([s.b for s in shards]), P.b)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = make_problem("small", seed=0)
    S = shard_rows(P, rank, world)
    L = O.OracleLogistic(S.indptr, S.indices, S.data, (S.n, S.d), S.b)
    rng = np.random.default_rng(3)
    x, v = rng.standard_normal(P.d), rng.standard_normal(P.d)
    # what every rank computes locally, then one SUM allreduce of the d-vector (and of the loss partial)
    part_hvp = torch.from_numpy(L.At_dot(L.hess_weights(x) * L.A_dot(v)))
    part_grad = torch.from_numpy(L.At_dot(O.expit(L.mat_vec_product(x)) - L.b))
    Ax = L.mat_vec_product(x)
    part_val = torch.tensor([np.sum((1 - L.b) * Ax - O.logsig(Ax))])
    for t in (part_hvp, part_grad, part_val):
        dist.all_reduce(t)
    if rank == 0:
        out.put((part_hvp.numpy() / P.n, part_grad.numpy() / P.n, float(part_val[0]) / P.n))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reduction_matches_single_shard():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    hvp, grad, val = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = make_problem("small", seed=0)
    L = O.OracleLogistic(P.indptr, P.indices, P.data, (P.n, P.d), P.b)
    rng = np.random.default_rng(3)
    x, v = rng.standard_normal(P.d), rng.standard_normal(P.d)
    assert np.linalg.norm(hvp - L.hess_vec_prod(x, v)) <= 1e-13 * np.linalg.norm(hvp)
    assert np.linalg.norm(grad - L.gradient(x)) <= 1e-13 * np.linalg.norm(grad)
    assert abs(val - L.value(x)) <= 1e-13 * abs(val)
