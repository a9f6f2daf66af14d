"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group (127.0.0.1), instance shards tile the batch
with no overlap, the job time is the max over ranks, throughput is total units / slowest rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parametron_b200 as pm


def test_instance_shard_partitions():
    for total in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            spans = [pm.shard.instance_shard(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            nxt = 0
            for first, count in spans:
                assert first == nxt or count == 0
                nxt = first + count
            assert max(c for _, c in spans) - min(c for _, c in spans if True) <= -(-total // world)
    with pytest.raises(ValueError):
        pm.shard.instance_shard(8, 2, 2)
    assert pm.shard.aggregate_throughput([4096, 4096], [1.0, 2.0]) == 4096.0


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = pm.shard.instance_shard(total, rank, world)
        counts = pm.shard.gather_counts(count)
        # pretend rank r needs (r + 1) seconds: the job takes the slowest rank's time
        t = pm.shard.max_over_ranks(float(rank + 1))
        firsts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(firsts, torch.tensor([first]))
        dist.barrier()
        q.put((rank, first, count, counts, t, [int(f.item()) for f in firsts]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_and_timing():
    world, total = 2, 8193
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, counts0, t0, firsts0), (r1, f1, c1, counts1, t1, firsts1) = res
    assert (f0, c0) == (0, 4097) and (f1, c1) == (4097, 4096)
    assert counts0 == counts1 == [4097, 4096] and firsts0 == firsts1 == [0, 4097]
    assert t0 == t1 == 2.0                                   # max over ranks, seen by both
    assert pm.shard.aggregate_throughput(counts0, [1.0, 2.0]) == total / 2.0
