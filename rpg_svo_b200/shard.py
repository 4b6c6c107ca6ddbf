"""Sharding of independent hot-path units (frame pairs, seeds, features) over ranks: SURVEY.md 8e.

The path has no exchange step: every unit is independent, so rank r of `world` simply owns its own synthetic camera
stream (weak scaling) or a contiguous slice of a fixed set of units, and the only collectives are the benchmark's barrier
and the max-over-ranks of the device time.  `RankGroup` is the one place that logic lives: bench.py drives it with the
NCCL backend on the GPU box, tests/test_shard_gloo.py drives the SAME class with gloo on CPU.
"""
from __future__ import annotations

import os


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of rank's units; sizes differ by at most one, all units covered once."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def stream_seed(rank: int) -> int:
    """Seed of the synthetic camera stream owned by a rank (one stream per GPU)."""
    return 1000 + rank


def aggregate_throughput(units_per_rank: int, world: int, seconds_max: float) -> float:
    """Whole-job throughput: units all ranks processed / max-over-ranks time (weak scaling)."""
    return units_per_rank * world / seconds_max


def _parse_cpulist(text: str) -> list[int]:
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def numa_node_of_gpu(pci_bus_id: str) -> int | None:
    """NUMA node the GPU's PCIe root hangs off (sysfs), None when unknown / single node."""
    pci = pci_bus_id.lower()
    if len(pci.split(":")[0]) == 8:  # CUDA prints an 8-digit domain, sysfs uses 4
        pci = pci[4:]
    try:
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
        return node if node >= 0 else None
    except (OSError, ValueError):
        return None


def bind_to_numa_node(node: int | None) -> dict:
    """Restrict this process (and therefore its pinned-memory allocations' first touch) to the CPUs of `node`,
    intersected with the CPUs the process may use.  Returns what was done."""
    info = {"numa_node": node, "bound": False}
    if node is None:
        return info
    try:
        cpus = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info.update(bound=True, cpus=len(use))
    except OSError as e:
        info["error"] = str(e)
    return info


def host_cpu_limits() -> dict:
    """What the CPU arm can really use: logical CPUs present, the affinity mask, and the cgroup quota."""
    out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            out["cgroup_" + os.path.basename(path)] = open(path).read().strip()
        except OSError:
            pass
    return out


def usable_threads() -> int:
    """Threads a CPU leg should start: the affinity mask, capped by a cgroup v2 quota when one is set."""
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


class RankGroup:
    """One process per GPU (or per CPU rank under gloo).  Weak scaling: every rank owns `units_per_rank` units of its
    own stream; timing = barrier, local device time, max over ranks."""

    def __init__(self, backend: str | None = None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend or "gloo", **kw)
            self.dist = dist

    @property
    def seed(self) -> int:
        return stream_seed(self.rank)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, value: float) -> float:
        """Max of a python float over the group (identity for a single process)."""
        if self.dist is None:
            return float(value)
        import torch

        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def throughput(self, units_per_rank: int, local_seconds: float) -> float:
        """Whole-job units/s: all ranks' units over the slowest rank's time."""
        return aggregate_throughput(units_per_rank, self.world, self.max_over_ranks(local_seconds))

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
