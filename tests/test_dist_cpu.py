"""world_size-2 gloo test of the multi-rank host logic bench.py relies on: contiguous channel
sharding with no data-path collective, and the max-over-ranks time reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from airspy_fmradion_b200.shard import channel_range


def test_channel_range_partition():
    for total in (0, 1, 7, 8, 256, 1001):
        for world in (1, 2, 3, 8):
            got = [channel_range(r, world, total) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            for a, b in zip(got, got[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        channel_range(2, 2, 8)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = channel_range(rank, world, 37)
    # each rank "processes" its own channels; the only exchanged values are counters and times
    t = torch.tensor([10.0 + 3.0 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, float(t.item()), int(n.item()), lo, hi))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 13.0      # max over ranks
    assert res[0][2] == res[1][2] == 37        # all channels covered exactly once
    assert res[0][4] == res[1][3]
