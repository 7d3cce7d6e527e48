"""Process-group plumbing for row-sharded runs (one process per GPU, torch.distributed).

torch.distributed only carries the 128-byte NCCL unique id from rank 0 to the other ranks; the data-path
collective (one allreduce of the d-vector per transpose sweep) is issued by libkcrn on its own communicator
and stream, inside the captured Lanczos graph.
"""
from __future__ import annotations

import numpy as np


def join_communicator(engine, rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    if world <= 1:
        return
    uid = engine.nccl_unique_id() if rank == 0 else bytes(128)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(np.frombuffer(uid, dtype=np.uint8).copy(), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    engine.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)


def shard_rows(P, rank: int, world: int):
    """Contiguous row block of problem P for this rank, balanced by non-zeros."""
    from .synth import balanced_row_cuts

    cuts = balanced_row_cuts(P.indptr, world)
    return P.row_slice(int(cuts[rank]), int(cuts[rank + 1]))
