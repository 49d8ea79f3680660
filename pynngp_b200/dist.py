"""Multi-GPU plumbing: one process per GPU (torchrun), contiguous shards of the ordering, one
allreduce of the 3*K partial statistics per evaluation (NCCL over NVLink on the box, gloo on CPU
in the tests).  Nothing here touches the data path: locations are independent given the replicated
coordinates, so the only exchange steps are
  (1) the one-off assembly of the neighbour table when stage 1 was split across ranks, and
  (2) the sum of {sum log F, sum r^2/F, n_bad}.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of the ordering owned by `rank` (equal counts; the likelihood's work
    per location is uniform)."""
    if not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    return (n * rank) // world, (n * (rank + 1)) // world


def knn_tile_split(rank: int, world: int):
    """Stage 1's work grows with i, so query tiles are dealt round-robin from the heavy end:
    (tile_offset, tile_stride) for nngp_build_neighbors."""
    return rank, world


def get_world(group=None):
    """(rank, world) of the active torch.distributed group, or (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
    except Exception:  # torch absent: single process
        return 0, 1
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def allreduce_stats(stats, group=None):
    """Sum a (K, 3) statistics array over ranks.  `stats` is a numpy array (host path / gloo) or a
    torch tensor (device path / NCCL, reduced in place with no host round trip)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    if isinstance(stats, torch.Tensor):
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        return stats
    t = torch.from_numpy(np.ascontiguousarray(stats, dtype=np.float64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def assemble_table_max(table, group=None):
    """Combine per-rank neighbour tables in which foreign rows hold ROW_UNSET (-2): elementwise MAX.
    `table` is an int32 torch tensor (device, in place) or numpy array."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return table
    if isinstance(table, torch.Tensor):
        dist.all_reduce(table, op=dist.ReduceOp.MAX, group=group)
        return table
    t = torch.from_numpy(np.ascontiguousarray(table, dtype=np.int32))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy()


class DevicePtrView:
    """Wraps a raw device pointer for torch.as_tensor(..., device='cuda') via
    __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
            "version": 3, "strides": None,
        }
