"""Sharding a sequence database over ranks (one process per GPU) and gathering the per-sequence scores.

The scan has no exchange step: sequences are independent and the model is replicated.  The only exchange is the gather
of fp32 scores at the end.  On GPUs it is FUSED into the scan (``FusedGather``: the kernel stores every score into all
ranks' copies of the gathered array over NVLink peer memory, no collective); ``gather_scores`` is the plain
all-gather (NCCL on GPUs; the same code runs over gloo on CPU in the tests).
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


class FusedGather:
    """Scan + gather in one kernel launch per rank (``msv_cuda_db_score_gather``).

    Every rank owns a *symmetric* buffer of ``world * slot`` floats (``torch.distributed._symmetric_memory``); the peers'
    buffers are mapped into this process over NVLink / NVSwitch.  ``scan(model, database)`` launches the MSV scan of this
    rank's shard with all ``world`` buffers as destinations -- the lane that finishes sequence q stores its score at
    ``rank * slot + q`` of every copy -- and then a device-side barrier on the current stream.  When that stream reaches
    the end of ``scan``, ``self.scores`` holds the whole job's scores on every rank (rank r's shard at
    ``[r * slot, r * slot + counts[r])``).  Needs one GPU per rank in one NVLink domain; raises whatever symmetric memory
    raises otherwise -- callers fall back to ``gather_scores``.
    """

    def __init__(self, slot: int, device, group=None, total: int | None = None, first_index: int | None = None) -> None:
        """Default layout: ``world`` slots of ``slot`` floats, rank r's shard at ``r * slot``.  With ``total`` and
        ``first_index`` the buffer is the job's score array itself (``total`` floats, global sequence order) and this
        rank's shard starts at ``first_index`` -- the layout of a cell-balanced cut of ONE database."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.slot = dist.get_world_size(group), dist.get_rank(group), max(int(slot), 1)
        if self.world > 8:
            raise ValueError("the fused gather addresses at most 8 GPUs (one NVSwitch domain)")
        if total is None:
            total, first_index = self.world * self.slot, self.rank * self.slot
        self.total, self.first_index = max(int(total), 1), int(first_index or 0)
        self.scores = symm_mem.empty(self.total, dtype=torch.float32, device=device)
        self.scores.fill_(float("nan"))
        self._handle = symm_mem.rendezvous(self.scores, group)
        ptrs = [int(p) for p in self._handle.buffer_ptrs]
        self._copies = [ptrs[self.rank]] + [ptrs[r] for r in range(self.world) if r != self.rank]  # own copy first

    def scan(self, model, database, stream: int = 0) -> None:
        """Resident database: one launch + the device-side barrier, asynchronous on the current stream."""
        database.score_gather(model, self._copies, self.first_index, stream)
        self._handle.barrier(channel=0)

    def scan_host(self, model, residues, offsets) -> None:
        """End to end: this rank's HOST buffers in (msv_cuda_score_batch_gather: upload, bucketing and scan pipelined,
        scores stored into every rank's copy), then the device-side barrier.  Synchronous for this rank's scan."""
        model.score_batch_gather(residues, offsets, self._copies, self.first_index)
        self._handle.barrier(channel=0)

    def shard(self, r: int, count: int):
        return self.scores[r * self.slot: r * self.slot + count]
