"""Batch sharding across the GPUs of one node (one process per GPU, torch.distributed / NCCL over NVLink).

The path has no exchange inside the model: samples are independent and weights are replicated.  The
only collective is the result aggregation the reference does with all_gather(list) + all_gather_object + per-meter
all_reduce (evaluate.py:185-222, utils/utils.py:176-198); here it is ONE all-gather of [B_local, n] fp32 per step.
"""
from __future__ import annotations

import torch


def shard_range(n_samples: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous ceil(n/world) samples per rank (the last shards may be short or empty)."""
    per = (n_samples + world - 1) // world
    lo = min(rank * per, n_samples)
    return lo, min(lo + per, n_samples)


def pad_shard(x: torch.Tensor, rows: int) -> torch.Tensor:
    """Pads a [b, n] shard with zero rows to `rows` so every rank contributes the same message size."""
    if x.shape[0] == rows:
        return x.contiguous()
    out = x.new_zeros((rows,) + tuple(x.shape[1:]))
    out[: x.shape[0]] = x
    return out


def gather_contacts(local: torch.Tensor, dist=None, n_samples: int | None = None) -> torch.Tensor:
    """[B_local, n] per rank -> [B_total, n] on every rank, rank-major (= sample order of shard_range).
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    rows = local.shape[0]
    if n_samples is not None:
        rows = (n_samples + world - 1) // world
        local = pad_shard(local, rows)
    out = local.new_empty((world * rows,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out if n_samples is None else out[:n_samples]
