"""Multi-GPU plumbing for the data-parallel hot path: utterance sharding and
the single weight broadcast at init.  torch.distributed is plumbing only; the
path has no per-step collective."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of utterances for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_bytes(blob, src: int = 0, device=None) -> bytes:
    """Replicates a byte string from rank `src` to every rank with ONE
    broadcast of the payload (plus an 8-byte length).  Works on the nccl
    backend (device="cuda:<local>") and on gloo (device=None / "cpu")."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(blob)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    rank = dist.get_rank()
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if rank == src:
        buf = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(dev)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src)
    return buf.cpu().numpy().tobytes()


def gather_token_lists(tokens: List[List[int]], dst: int = 0):
    """Collects per-rank results on `dst` (host-side, optional: results are
    ~100 ids per utterance)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [tokens]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(tokens, out, dst=dst)
    return out
