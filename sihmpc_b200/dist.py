# -*- coding: utf-8 -*-
"""Multi-GPU sharding of a scenario set: one process per GPU (torchrun), scenarios are
independent units, so rank r of G owns the contiguous index range [r*B/G, (r+1)*B/G) and NO
data-path collective exists during stepping.  The only exchange is the reduction / gather of the
closed-loop statistics at the end (NCCL over NVLink on GPUs; gloo in the CPU tests).

SURVEY.md section 8(e).  The reference is single-process (no collective call site anywhere).
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split; the first (total % world) ranks get one extra scenario."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend)
    return rank, world


def reduce_moments(agg: torch.Tensor) -> torch.Tensor:
    """agg = [3][nstat] (sum, sum of squares, max) per rank -> the same over all ranks, in place.
    ONE collective: an all-gather of the 3 x nstat block (a few hundred bytes per rank), then sum / sum / max
    over the rank axis locally -- the payload is latency, not bandwidth, so one launch instead of a SUM
    all-reduce plus a MAX all-reduce."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        world = dist.get_world_size()
        flat = agg.contiguous().reshape(-1)
        out = torch.empty(world * flat.numel(), dtype=agg.dtype, device=agg.device)
        dist.all_gather_into_tensor(out, flat)
        out = out.reshape(world, 3, -1)
        agg[:2].copy_(out[:, :2].sum(dim=0))
        agg[2].copy_(out[:, 2].max(dim=0).values)
    return agg


def gather_stats(stats: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather of the per-scenario statistics block [local][nstat] into [total][nstat]
    (config 5: 1M x 7 FP64 = 56 MB over 8 ranks).  Uneven shards are padded to the largest one."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, stats.shape[1], dtype=stats.dtype, device=stats.device)
    pad[:stats.shape[0]] = stats
    out = torch.empty(world * mx, stats.shape[1], dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def moments_from_stats(stats: np.ndarray) -> np.ndarray:
    """Host mirror of sihmpc_stats_reduce (used by the CPU tests of the exchange step)."""
    s = np.asarray(stats, dtype=np.float64)
    return np.stack([s.sum(axis=0), (s * s).sum(axis=0), s.max(axis=0)])
