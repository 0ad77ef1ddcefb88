"""CPU: the N > 1 host logic (stream partition + counter aggregation) with world_size 2 over gloo."""
import os
import socket

import pytest

from rsvio_b200 import sharding


def test_partition_is_a_disjoint_cover():
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.partition_streams(total, world, r) for r in range(world)]
            flat = [s for p in parts for s in p]
            assert flat == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sharding.weak_streams(64, 3) == list(range(192, 256))
    with pytest.raises(ValueError):
        sharding.partition_streams(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.rank_info() == (rank, rank, world)
        mine = sharding.partition_streams(5, world, rank)
        # every stream contributes 100 tracks and 1 frame per step, 3 steps; rank 1 is slower
        tracks, frames, iters, ms = sharding.aggregate(dist, torch.device("cpu"), 300 * len(mine), 3 * len(mine),
                                                       7 * len(mine), 10.0 + 5.0 * rank)
        q.put((rank, mine, tracks, frames, iters, ms))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_aggregate_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out[0][1] == [0, 1, 2] and out[1][1] == [3, 4]
    for _, _, tracks, frames, iters, ms in out:
        assert (tracks, frames, iters, ms) == (1500.0, 15.0, 35.0, 15.0)   # sums over ranks, max time
    tps, fps = sharding.throughput(1500.0, 15.0, 15.0)
    assert abs(tps - 1e5) < 1e-6 and abs(fps - 1000.0) < 1e-9


def test_aggregate_single_process():
    assert sharding.aggregate(None, None, 10, 2, 5, 4.0) == (10.0, 2.0, 5.0, 4.0)


def test_bench_reads_every_live_record():
    """bench.sum_live_records (the e2e loop's consumer of the host-mapped result ring) equals the per-group loop it
    replaced: every record k < counts[group] of every group, nothing beyond"""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    rng = np.random.default_rng(7)
    dt = np.dtype([("id", "<u8"), ("x", "<f4"), ("y", "<f4")])
    rec = np.zeros((12, 40), dt)
    rec["x"] = rng.uniform(0, 752, rec.shape)
    rec["y"] = rng.uniform(0, 480, rec.shape)
    for cnt in (np.array([0] * 12, np.int32), rng.integers(0, 41, 12).astype(np.int32), np.full(12, 40, np.int32)):
        want = sum(float(rec["x"][g, :c].astype(np.float64).sum()) + float(rec["y"][g, :c].astype(np.float64).sum())
                   for g, c in enumerate(cnt))
        assert abs(bench.sum_live_records(rec, cnt) - want) <= 1e-6 * max(1.0, abs(want))
