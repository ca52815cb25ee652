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
