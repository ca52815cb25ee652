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


def gather_parts(part, dst: int = 0):
    """collect the per-rank pat parts (bytes, or a list of (chromosome order, bytes)) on `dst` in rank order (the `cat` of
    bam2pat.py:408); other ranks get None"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [part]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(part, out, dst=dst)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# running the CLIs under torchrun: `python -m torch.distributed.run --nproc-per-node N -m wgbs_tools_b200.cli <cmd> ...`
# ----------------------------------------------------------------------------------------------------------------------
def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank).  Under torchrun (WORLD_SIZE > 1) the process group is created once: NCCL over
    NVLink / NVSwitch by default, gloo when WGBS_DIST_BACKEND=gloo (CPU tests).  Alone: (0, 1, 0), nothing is initialised."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        backend = backend or os.environ.get("WGBS_DIST_BACKEND", "nccl")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def world() -> tuple[int, int]:
    """(rank, world) of the initialised process group, (0, 1) without one"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_lines(text: bytes, rank: int, world: int) -> bytes:
    """the rank-th of `world` contiguous pieces of a text, cut at line boundaries near equal byte offsets.  Record-sharding
    for the steps whose result is a SUM over records (pat2beta counts, homog bins): every rank sees all blocks / sites."""
    if world <= 1:
        return text
    n = len(text)

    def cut(k: int) -> int:
        if k <= 0:
            return 0
        if k >= world:
            return n
        p = text.find(b"\n", (n * k) // world)
        return n if p < 0 else p + 1
    return text[cut(rank):cut(rank + 1)]


def reduce_np(a: np.ndarray, dst: int = 0) -> np.ndarray:
    """sum a host array over the ranks onto `dst` (through a CUDA tensor when the backend is NCCL); other ranks get their
    own array back.  A no-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return a
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t.cpu().numpy() if dist.get_rank() == dst else a


class DistSolver:
    """segment: the independent DPs of one round (chunks, or the stitching patches of one tree level) are dealt
    round-robin to the ranks; every rank solves its share on its own GPU and all ranks receive all borders
    (a few KB per DP), so the deterministic host stitching continues identically everywhere."""

    def __init__(self, solve):
        self.solve = solve

    def __call__(self, sites):
        import torch.distributed as dist
        rank, n = world()
        if n == 1:
            return self.solve(sites)
        mine = list(sites[rank::n])
        res = self.solve(mine) if mine else []
        allres = [None] * n
        dist.all_gather_object(allres, [np.asarray(r) for r in res])
        out = [None] * len(sites)
        for r in range(n):
            for j, b in enumerate(allres[r]):
                out[r + j * n] = b
        return out
