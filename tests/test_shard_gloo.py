"""CPU, world_size 2, gloo: the N>1 path of the benchmark (independent units sharded over ranks, no
data-path collective; barrier + max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rpg_svo_b200 import shard


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 256, 2000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                b, e = shard.shard_range(n, r, world)
                assert e - b in (n // world, n // world + 1)
                got += list(range(b, e))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard.shard_range(256, rank, world)          # BASELINE config C4: 256 pairs over the ranks
    my_ms = 10.0 + 5.0 * rank                            # pretend device time of this rank
    dist.barrier()
    worst = shard.max_over_ranks(my_ms, dist)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([e - b]))
    q.put((rank, worst, int(sum(c.item() for c in counts)), shard.stream_seed(rank)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [15.0, 15.0]           # every rank sees the max
    assert [r[2] for r in res] == [256, 256]             # all units owned exactly once
    assert [r[3] for r in res] == [1000, 1001]           # one synthetic stream per rank
    assert shard.aggregate_throughput(128, 2, 0.015) == pytest.approx(256 / 0.015)
