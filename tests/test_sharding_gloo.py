"""world_size-2 gloo tests (CPU) of the multi-rank host logic: env partitioning, max-over-ranks
timing, whole-job throughput aggregation."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from simfire_b200.sharding import env_shard, row_slabs


def test_env_shard_partitions():
    for total in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            parts = [env_shard(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == total
            for (a, n), (b, _) in zip(parts, parts[1:]):
                assert a + n == b
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    with pytest.raises(ValueError):
        env_shard(8, 2, 2)
    assert row_slabs(8192, 8) == [(1024 * r, 1024) for r in range(8)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))  # fmt: skip
    from simfire_b200.sharding import RankContext, aggregate_throughput

    ctx = RankContext.from_env(backend="gloo")
    first, n = env_shard(1023, ctx.world, ctx.rank)
    ctx.barrier()
    ctx.barrier(host=True)  # the bench's bracket around its timed region (on an NCCL context: a gloo side group)
    ms = ctx.max(10.0 + 5.0 * rank)  # rank 1 is slower
    total = ctx.sum(n)
    thr = aggregate_throughput(ctx, cells_this_rank=n * 100, steps=4, seconds_this_rank=0.5 * (rank + 1))
    q.put((rank, first, n, ms, total, thr))
    ctx.close()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, n0, ms0, t0, thr0), (r1, f1, n1, ms1, t1, thr1) = res
    assert (f0, n0, f1, n1) == (0, 512, 512, 511)
    assert ms0 == ms1 == 15.0  # the slowest rank's time on every rank
    assert t0 == t1 == 1023
    assert thr0 == thr1 == pytest.approx(1023 * 100 * 4 / 1.0)
