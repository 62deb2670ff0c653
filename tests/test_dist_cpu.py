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


def _transport_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from airspy_fmradion_b200.shard import SingleHomedTransport
    total, row = 5, 12  # uneven: 3 + 2 channels, padded to 3 rows per rank
    tr = SingleHomedTransport(total, root=0)
    full = None
    if rank == 0:
        full = (torch.arange(total * row, dtype=torch.int32).reshape(total, row) % 251).to(torch.uint8)
    shard = tr.scatter_rows(full, row, torch.uint8, torch.device("cpu"))
    # stand-in for the decode: every row -> its running sum as int16 (the real one needs a GPU)
    local = shard.to(torch.int16).cumsum(dim=1).to(torch.int16)
    out = tr.gather_rows(local)
    ok = True
    if rank == 0:
        want = full.to(torch.int16).cumsum(dim=1).to(torch.int16)
        ok = out.shape == (total, row) and bool((out == want).all())
    else:
        ok = out is None
    q.put((rank, tr.lo, tr.hi, tr.rows_per_rank, tuple(shard.shape), bool(ok),
           bool((shard[tr.n_local:] == 0).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_single_homed_transport_two_rank_gloo():
    """Ingest on rank 0: rows scattered in the file's byte format, results gathered back (uneven partition)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_transport_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[3] == 3 and r[4] == (3, 12) and r[5] and r[6] for r in res)
