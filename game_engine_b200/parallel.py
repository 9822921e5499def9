"""Session-level data parallelism: contiguous session-id shards and the one exchange step of the path.

Sessions are independent (state, Philox counter and table are per session / read-only), so the path shards
with no data-path collective; the only exchange is a final SUM all-reduce of the u64 statistics words
(win rate + phase-length histogram, SPEC.md section 6) — NCCL over NVLink on GPUs, gloo in the CPU tests.
Integer sums are order independent, so the reduced result is identical for any world size.
The reference has no analogue (one asyncio loop per room).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from .capi import STATS_LEN


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of the contiguous session range [rank*n/world, (rank+1)*n/world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi - lo


class _CudaArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def stats_tensor(batch, device) -> torch.Tensor:
    """Zero-copy int64 view of a batch's device statistics snapshot (after batch.stats_refresh())."""
    return torch.as_tensor(_CudaArray(batch.stats_device_ptr(), STATS_LEN), device=device)


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """SUM all-reduce in place when a process group exists; counts stay below 2^63 so int64 is exact."""
    if stats.dtype != torch.int64:
        raise TypeError("statistics are exchanged as int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def win_rates(stats) -> dict:
    v, w = int(stats[2]), int(stats[3])
    tot = max(1, v + w)
    return {"villagers": v / tot, "werewolves": w / tot, "finished": v + w, "unfinished": int(stats[1])}
