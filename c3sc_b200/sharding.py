"""Fiber-batch sharding across the GPUs of one box.

Fibers are independent given the FT cores: rank g owns the contiguous block
[g*ceil(F/G), min(F,(g+1)*ceil(F/G))) -- a STABLE map, so policy-evaluation rows stay on
the GPU that produced them.  The only exchange is a broadcast of the cores when they
change and an all-gather of the backed-up values.  Works with any torch.distributed
backend (nccl on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_size(F: int, world: int) -> int:
    return (F + world - 1) // world


def shard_range(F: int, world: int, rank: int) -> tuple[int, int]:
    per = shard_size(F, world)
    lo = min(F, rank * per)
    return lo, min(F, lo + per)


def shard_fibers(dim_vary: np.ndarray, fixed_ind: np.ndarray, world: int, rank: int):
    """This rank's block, padded (by repeating its last fiber, or fiber 0 of the batch when the
    block is empty) to the common shard size so the all-gather is regular."""
    F = int(dim_vary.shape[0])
    per = shard_size(F, world)
    lo, hi = shard_range(F, world, rank)
    idx = np.arange(lo, hi)
    if idx.size < per:
        fill = idx[-1] if idx.size else 0
        idx = np.concatenate([idx, np.full(per - idx.size, fill, dtype=idx.dtype)])
    return dim_vary[idx].copy(), fixed_ind[idx].copy(), hi - lo


def broadcast_cores(flat_cores, src: int = 0, group=None):
    """In-place broadcast of the contiguous core buffer (a torch tensor) from `src`."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_cores, src=src, group=group)
    return flat_cores


def gather_values(local_values, F: int, ldo: int, out=None, group=None):
    """All-gather the per-rank [per*ldo] value blocks and return the first F*ldo entries in
    global fiber order (a view of `out`)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_values[: F * ldo]
    per = local_values.numel() // ldo
    if out is None:
        out = torch.empty(world * per * ldo, dtype=local_values.dtype, device=local_values.device)
    dist.all_gather_into_tensor(out, local_values, group=group)
    return out[: F * ldo]
