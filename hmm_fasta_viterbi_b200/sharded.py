"""Sharding a sequence database over ranks (one process per GPU) and gathering the per-sequence scores.

The scan has no exchange step: sequences are independent and the model is replicated.  The only collective is the
gather of fp32 scores at the end (NCCL over NVLink on GPUs; the same code runs over gloo on CPU in the tests).
Slices are contiguous and balanced by residue count (= DP cell count for a fixed model), computed by the C ABI helper
``msv_host_partition_by_cells``.
"""
from __future__ import annotations

import numpy as np

from . import _cabi


def shard_bounds(offsets: np.ndarray, world: int) -> np.ndarray:
    """world+1 sequence indices; rank r owns sequences [bounds[r], bounds[r+1])."""
    return _cabi.partition_by_cells(np.ascontiguousarray(offsets, np.uint64), world)


def local_slice(residues: np.ndarray, offsets: np.ndarray, rank: int, world: int):
    """This rank's contiguous slice as (residues, offsets rebased to 0, first sequence index, one-past-last)."""
    bounds = shard_bounds(offsets, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    base = int(offsets[lo])
    sl_res = residues[base:int(offsets[hi])]
    sl_off = (np.asarray(offsets[lo:hi + 1], dtype=np.uint64) - np.uint64(base)).astype(np.uint64)
    return sl_res, sl_off, lo, hi


def gather_scores(local_scores, counts, group=None):
    """All-gather per-rank score tensors of different lengths; returns the scores of the whole database in global
    order on every rank.  `local_scores` is a 1-D float32 torch tensor (CUDA for NCCL, CPU for gloo) holding at least
    counts[rank] valid entries; `counts` lists every rank's sequence count."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    width = max(max(counts), 1)
    padded = torch.zeros(width, dtype=torch.float32, device=local_scores.device)
    padded[: counts[rank]] = local_scores[: counts[rank]]
    gathered = torch.empty(world * width, dtype=torch.float32, device=local_scores.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    return torch.cat([gathered[r * width: r * width + counts[r]] for r in range(world)])
