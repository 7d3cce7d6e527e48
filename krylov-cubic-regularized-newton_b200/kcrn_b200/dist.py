"""Process-group plumbing for row-sharded runs (one process per GPU, torch.distributed).

torch.distributed only carries the 128-byte NCCL unique id from rank 0 to the other ranks and the CUDA IPC
handles of the peer exchange buffers; the data-path collective (one sum of the d-vector per transpose sweep) is
done by libkcrn itself inside the captured Lanczos graph -- fused into the Lanczos epilogue kernel over NVLink
peer memory when d is small, an NCCL allreduce on its own communicator otherwise.
"""
from __future__ import annotations

import numpy as np


def join_communicator(engine, rank: int, world: int, peer_exchange: bool = True) -> None:
    import torch
    import torch.distributed as dist

    if world <= 1:
        return
    uid = engine.nccl_unique_id() if rank == 0 else bytes(128)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(np.frombuffer(uid, dtype=np.uint8).copy(), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    engine.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)
    # NVLink peer exchange buffers for the fused Lanczos step (CUDA IPC handles, all-gathered in rank order): only
    # when every rank owns a different GPU of this host and all of them can map each other's memory -- the
    # ranks spin on each other's flags inside a kernel, which two processes sharing one GPU could not survive
    if not peer_exchange or world > 16 or dev != "cuda":
        return
    import socket
    import zlib

    me = torch.tensor([zlib.crc32(socket.gethostname().encode()), torch.cuda.current_device()],
                      dtype=torch.int64, device=dev)
    every = [torch.empty_like(me) for _ in range(world)]
    dist.all_gather(every, me)
    every = [tuple(int(v) for v in e.cpu()) for e in every]
    ok = len({e[0] for e in every}) == 1 and len({e[1] for e in every}) == world
    ok = ok and all(k == every[rank][1] or torch.cuda.can_device_access_peer(every[rank][1], k)
                    for _, k in every)
    agreed = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
    if int(agreed.item()) == 0:
        return
    mine = torch.tensor(np.frombuffer(engine.p2p_export(), dtype=np.uint8).copy(), dtype=torch.uint8, device=dev)
    blobs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(blobs, mine)
    engine.p2p_attach(b"".join(bytes(e.cpu().numpy().tobytes()) for e in blobs))
    dist.barrier()


def shard_rows(P, rank: int, world: int):
    """Contiguous row block of problem P for this rank, balanced by non-zeros.

    Labels: build every shard's oracle with ``LogisticRegression(shard, shard_b, label_encoding=label_encoding(P.b))``
    (optimizer.loss) -- the {0, 1} encoding is a property of the GLOBAL label vector (reference loss.py:190-207)."""
    from .synth import balanced_row_cuts

    cuts = balanced_row_cuts(P.indptr, world)
    return P.row_slice(int(cuts[rank]), int(cuts[rank + 1]))
