"""Multi-GPU plumbing (one process per GPU, torch.distributed): shard assignment and the single exchange step.

The hot path shards the way the reference parallelises (one worker per chromosome / region, bam2pat.py:319-346): shards
are independent, pat parts are concatenated in chromosome order, and the per-shard `int32[N,2]` beta counts are summed
with ONE reduce over NCCL/NVLink *before* the non-linear uint8 trim.  No other collective exists on this
path."""
from __future__ import annotations

import numpy as np


def lpt_assign(weights, world: int) -> list[list[int]]:
    """Longest-processing-time-first bin packing of shards (e.g. chromosomes weighted by read count) onto `world` ranks.
    Returns, per rank, the shard ids it owns, each list in ascending shard id (= chromosome order)."""
    w = np.asarray(weights, dtype=np.float64)
    loads = np.zeros(world)
    owner = [[] for _ in range(world)]
    for i in np.argsort(-w, kind="stable").tolist():
        r = int(np.argmin(loads))
        owner[r].append(i); loads[r] += w[i]
    return [sorted(o) for o in owner]


def split_reads_evenly(n_lines: int, world: int) -> list[tuple[int, int]]:
    """contiguous [begin, end) line ranges, one per rank (read-sharded pileup of a single chromosome)"""
    per = (n_lines + world - 1) // world
    return [(min(r * per, n_lines), min((r + 1) * per, n_lines)) for r in range(world)]


def reduce_counts(counts, dst: int = 0):
    """sum the per-rank int32[N,2] (meth, cover) arrays onto `dst` (in place).  counts: torch tensor (CUDA -> NCCL over
    NVLink; CPU -> gloo in tests).  A no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM)
    return counts


def gather_parts(part: bytes, dst: int = 0):
    """collect the per-rank pat text parts on `dst` in rank order (the `cat` of bam2pat.py:408)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [part]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(part, out, dst=dst)
    return out
