"""CPU, world_size 2, gloo: the N>1 path of the benchmark.  bench.py's rank logic lives in rpg_svo_b200.shard.RankGroup
(rank / seed assignment, barrier, max-over-ranks timing, whole-job throughput, NUMA binding helpers); this test drives that
same class with the gloo backend -- independent units sharded over ranks, no data-path collective."""
import os
import socket

import pytest
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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    grp = shard.RankGroup("gloo")                        # what bench.py builds with "nccl"
    assert (grp.rank, grp.world) == (rank, world)
    my_ms = 10.0 + 5.0 * rank                            # pretend device time of this rank
    grp.barrier()
    worst = grp.max_over_ranks(my_ms)
    value = grp.throughput(32, my_ms * 1e-3)             # BASELINE configs[4]: 32 pairs per rank, weak scaling
    q.put((rank, worst, value, grp.seed))
    grp.close()


def test_two_rank_gloo_timing_and_throughput():
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
    assert [r[1] for r in res] == [15.0, 15.0]                       # every rank sees the max
    assert [r[2] for r in res] == [pytest.approx(64 / 0.015)] * 2    # all ranks' units over the slowest rank's time
    assert [r[3] for r in res] == [1000, 1001]                       # one synthetic stream per rank


def test_single_process_group_is_the_identity():
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    grp = shard.RankGroup("gloo")
    assert (grp.rank, grp.world, grp.seed) == (0, 1, 1000)
    grp.barrier()
    assert grp.max_over_ranks(3.5) == 3.5 and grp.throughput(100, 0.5) == 200.0
    grp.close()


def test_host_limits_and_numa_helpers():
    lim = shard.host_cpu_limits()
    assert lim["affinity"] >= 1 and 1 <= shard.usable_threads() <= lim["affinity"]
    assert shard._parse_cpulist("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11]
    assert shard.numa_node_of_gpu("0000:ff:1f.0") in (None, 0, 1, 2, 3)   # unknown device -> None
    assert shard.bind_to_numa_node(None) == {"numa_node": None, "bound": False}
