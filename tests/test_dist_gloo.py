"""world_size-2 gloo test (CPU) of the N>1 host path: shard assignment, per-rank counts, ONE reduce(sum) before the
uint8 trim, parts gathered in rank order.  The per-shard compute is the oracle port here (no GPU in this container); the
`-m gpu` suite covers the kernels, bench.py --gpus N the NCCL path."""
import os
import socket

import numpy as np
import pytest

from wgbs_tools_b200 import dist as wd
from wgbs_tools_b200 import synth


def test_lpt_assign_balances_and_keeps_order():
    w = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57, 1]
    for world in (1, 2, 4, 8):
        own = wd.lpt_assign(w, world)
        assert sorted(i for o in own for i in o) == list(range(len(w)))
        loads = [sum(w[i] for i in o) for o in own]
        assert max(loads) <= sum(w) / world * 1.15 + max(w) * (world > 4)
        assert all(o == sorted(o) for o in own)
    assert wd.split_reads_evenly(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import harness as H
    N = 4000
    idx, pats, cnt = synth.make_pat_records(3, 20_000, N, mean_len=6)
    cnt = cnt * 7                                                  # cover > 255 somewhere: the trim must come AFTER the reduce
    lines = synth.pat_text("chr1", idx, pats, cnt).splitlines(keepends=True)
    b, e = wd.split_reads_evenly(len(lines), world)[rank]
    mine = b"".join(lines[b:e])
    counts = torch.from_numpy(H.port_pat2beta(mine, 1, N + 1).copy())
    wd.reduce_counts(counts, 0)
    parts = wd.gather_parts(mine, 0)
    if rank == 0:
        full = H.port_pat2beta(b"".join(lines), 1, N + 1)
        assert np.array_equal(counts.numpy(), full)
        beta = H.port_trim(counts.numpy())
        assert beta.tobytes() == H.port_trim(full).tobytes()
        wrong = sum(H.port_trim(H.port_pat2beta(b"".join(lines[x:y]), 1, N + 1)).astype(np.int64) for x, y in wd.split_reads_evenly(len(lines), world))
        assert not np.array_equal(wrong, beta.astype(np.int64))     # trimming per shard first would be wrong
        assert b"".join(parts) == b"".join(lines)
        open(os.path.join(tmp, "ok"), "w").write("1")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_then_trim(oracle, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_shard_lines_partitions_at_line_boundaries():
    txt = b"".join(b"l%d\t%s\n" % (i, b"x" * (i % 17)) for i in range(1000)) + b"last-without-newline"
    for world in (1, 2, 3, 8, 50):
        parts = [wd.shard_lines(txt, r, world) for r in range(world)]
        assert b"".join(parts) == txt
        assert all(p.endswith(b"\n") for p in parts[:-1] if p)
    assert wd.shard_lines(b"", 0, 4) == b"" and [wd.shard_lines(b"a\n", r, 4) for r in range(4)].count(b"a\n") == 1


def _worker2(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import harness as H
    from wgbs_tools_b200 import segment as sg
    # homog: records sharded by line ranges, every rank sees all blocks, ONE reduce of the bins
    N = 5000
    idx, pats, cnt = synth.make_pat_records(11, 15_000, N, mean_len=6)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    blocks = synth.make_blocks(4, 1, N, mean_len=9.0)[:-40]                 # records past the last block exist: the cut-off rule is per record
    edges = np.array([0, 0.334, 0.667, 1], np.float32)
    mine = H.port_homog(wd.shard_lines(txt, rank, world), blocks, edges, 3)
    tot = wd.reduce_np(mine.astype(np.int32), 0)
    # segment: the DPs of a round dealt round-robin, all borders on all ranks, identical stitching everywhere
    betas = synth.make_betas(9, 3, 3000)
    loci = synth.make_genome(2, "chr1", 600_000, with_bases=False).loci[:3000].astype(np.int64)
    calls = []

    def solve(sites):
        calls.append(len(sites))
        return [H.port_segment([b[s - 1:e - 1] for b in betas], loci[s - 1:e - 1], 100, 1500, 15) + s for s, e in sites]
    got = sg.segment_regions([(1, 1501), (1501, 3001)], wd.DistSolver(solve), 400)
    every = [None] * world
    dist.all_gather_object(every, got.tolist())
    if rank == 0:
        assert np.array_equal(tot, H.port_homog(txt, blocks, edges, 3)) and tot.sum() > 1000
        ref_calls = []

        def solve1(sites):
            ref_calls.append(len(sites))
            return [H.port_segment([b[s - 1:e - 1] for b in betas], loci[s - 1:e - 1], 100, 1500, 15) + s for s, e in sites]
        exp = sg.segment_regions([(1, 1501), (1501, 3001)], solve1, 400)
        assert np.array_equal(got, exp) and all(e == exp.tolist() for e in every)
        assert sum(calls) < sum(ref_calls)                                   # this rank solved only its share
        open(os.path.join(tmp, "ok2"), "w").write("1")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_homog_reduce_and_segment_round_robin(oracle, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker2, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok2").exists()
