"""Multi-GPU plumbing of the reconstruction path: none of it touches pixel data.

Images are independent and blocks never exchange data (replication upsampling
stays inside an MCU, gid-decoding_jpg.adb:984-1007), so the path shards with NO
data-path collective: image i of a batch goes to GPU ``i mod G``.
The only cross-rank operations are the timing barrier and the max-over-ranks
reduction of the device time, which work the same over NCCL (GPUs) and gloo (CPU
tests).
"""
from __future__ import annotations

from typing import Sequence


def shard(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the images rank ``rank`` of ``world`` reconstructs (round-robin)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of world {world}")
    return list(range(rank, n_items, world))


def owner(index: int, world: int) -> int:
    return index % world


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of one float (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(units_per_rank: Sequence[float] | float, world: int, seconds_max: float) -> float:
    """Whole-job throughput: units processed by ALL ranks / the slowest rank's time."""
    total = units_per_rank * world if isinstance(units_per_rank, (int, float)) else sum(units_per_rank)
    return total / seconds_max
