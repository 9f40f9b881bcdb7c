"""world_size-2 gloo tests (CPU) of the path's only exchange step: observer statistics and scale
gradients combined across ranks must equal the single-process result on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _stats_of(t):
    rows = t.reshape(t.shape[0], -1) if t.dim() > 1 else t.reshape(1, -1)
    return torch.stack([rows.min(1)[0], rows.max(1)[0], rows.abs().max(1)[0], rows.abs().sum(1)], 1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dlmc_quant_b200 import dist as qdist
        gen = torch.Generator().manual_seed(2333)
        full = torch.randn(8, 6, 5, generator=gen)                 # channels x (batch*inner); batch split over ranks
        shard = full[:, rank * 3:(rank + 1) * 3]
        merged = qdist.sync_stats(_stats_of(shard))
        ok_stats = torch.allclose(merged, _stats_of(full), rtol=1e-6) and torch.equal(merged[:, :3], _stats_of(full)[:, :3])
        bad = shard.clone()
        if rank == 1:
            bad[2, 0, 0] = float("nan")
        m2 = qdist.sync_stats(_stats_of(bad).nan_to_num(nan=0.0) if False else _nan_stats(bad))
        ok_nan = bool(torch.isnan(m2[2]).all()) and not bool(torch.isnan(m2[[0, 1, 3]]).any())
        sse, rows = qdist.sync_sse(torch.full((80,), float(rank + 1)), 12.0)
        ok_sse = bool((sse == 3.0).all()) and rows == 24.0
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        qdist.allreduce_grads_(flat)
        ok_sum = torch.equal(flat, torch.arange(10, dtype=torch.float32) * 3)
        flat2 = torch.ones(4) * (rank + 1)
        qdist.allreduce_grads_(flat2, average=True)
        ok_avg = torch.equal(flat2, torch.full((4,), 1.5))
        qdist.set_enabled(False)
        ok_off = qdist.world_size() == 1 and qdist.sync_stats(merged) is merged
        qdist.set_enabled(True)
        q.put((rank, ok_stats, ok_nan, ok_sse, ok_sum, ok_avg, ok_off))
    finally:
        dist.destroy_process_group()


def _nan_stats(t):
    """what the statistics kernel reports for a row containing NaN: all four entries NaN."""
    s = _stats_of(t)
    nan = torch.isnan(t.reshape(t.shape[0], -1)).any(1)
    s[nan] = float("nan")
    return s


def test_stat_and_grad_exchange_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert all(r[1:]), f"rank {r[0]} failed: stats/nan/sse/sum/avg/off = {r[1:]}"


def test_single_process_is_identity():
    from dlmc_quant_b200 import dist as qdist
    s = torch.randn(4, 4)
    assert qdist.world_size() == 1 and qdist.sync_stats(s) is s
    sse, rows = qdist.sync_sse(s, 3.0)
    assert sse is s and rows == 3.0


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dlmc_quant_b200 import dist as qdist
        gen = torch.Generator().manual_seed(2333)
        calls = []

        def fn(block):              # stand-in for a per-channel observer: (scale, offset) per row
            calls.append(block.shape[0])
            return block.abs().max(1)[0] / 7, block.min(1)[0]
        oks = []
        for channels in (131, 128, 3):          # uneven split, even split, too small to shard
            rows = torch.randn(channels, 37, generator=gen)
            calls.clear()
            s, o = qdist.rows_sharded(rows, fn, min_rows_per_rank=16)
            ws, wo = rows.abs().max(1)[0] / 7, rows.min(1)[0]
            start, stop = qdist.row_block(channels, rank, world)
            sharded = channels >= world * 16
            oks.append(torch.equal(s, ws) and torch.equal(o, wo) and calls == [stop - start if sharded else channels])
        q.put((rank, *oks))
    finally:
        dist.destroy_process_group()


def test_per_channel_observer_rows_are_sharded_and_gathered_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert all(r[1:]), f"rank {r[0]}: uneven/even/small = {r[1:]}"


def test_row_block_partition_covers_all_rows():
    from dlmc_quant_b200 import dist as qdist
    for channels, world in ((131, 2), (27560, 8), (5, 8), (64, 4)):
        blocks = [qdist.row_block(channels, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == channels
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
