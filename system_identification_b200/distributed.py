"""Sample-parallel sharding of the path: each rank builds the Gram statistics of its contiguous time shard,
one SUM all-reduce of c^2 + c + 2 doubles (191 KB at c = 154) merges them, the solve runs on rank 0.
The reference has no parallelism of any kind (single Python thread); see SURVEY.md section 8e.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(N, rank, world_size):
    """Contiguous shard [lo, hi) of N samples for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(N), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks of the packed [G | r | s | n] statistics (NCCL for CUDA tensors, gloo for CPU)."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def broadcast_solution(x: torch.Tensor, src=0) -> torch.Tensor:
    rank, ws = world()
    if ws > 1:
        dist.broadcast(x, src=src)
    return x
