"""Multi-GPU driver logic: independent random restarts / staircase ranks, one per rank, no
collective on the data path; at the end the best certified solution is gathered.

`gather_best(...)` is the host-side protocol: all-gather of {f, certified}, the winner rule of the
C-ABI (`cora_b200_select_best`), broadcast of the winner's iterate.  With a NCCL communicator created
through the library (`native=True`, GPU ranks) the exchange runs inside `cora_b200_gather_best`
(ncclAllGather + ncclBroadcast over NVLink); otherwise it runs over `torch.distributed` (any backend --
the world_size-2 `gloo` CPU tests use this path, same rule, same result).
"""
from __future__ import annotations

import numpy as np

from . import capi


def restart_seed(rank: int, base: int = 0) -> int:
    """Restart g runs on rank g with seed base + g (BASELINE configs[3]: seeds 0..7)."""
    return base + rank


def make_native_comm(dist, device: int):
    """Create the library's own NCCL communicator: rank 0 makes the unique id, torch.distributed ships it."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return capi.NcclComm(device, world, rank, box[0])


def gather_best(dist, f: float, certified: bool, X: np.ndarray, handle=None, comm=None):
    """Returns (winner_rank, winner_f, X_winner) on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if handle is not None and comm is not None:
        return handle.gather_best(comm, world, rank, f, certified, X)
    import torch
    rec = torch.tensor([float(f), 1.0 if certified else 0.0], dtype=torch.float64)
    out = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(out, rec)
    fs = [float(o[0]) for o in out]
    cs = [int(o[1] != 0) for o in out]
    win = capi.select_best(fs, cs)
    buf = torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float64).T))  # column-major N x r as r x N
    dist.broadcast(buf, src=win)
    return win, fs[win], np.asfortranarray(buf.numpy().T)
