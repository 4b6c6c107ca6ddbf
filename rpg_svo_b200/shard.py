"""Sharding of independent hot-path units (frame pairs, seeds, features) over ranks: SURVEY.md 8e.

The path has no exchange step: every unit is independent, so rank r of `world` simply owns a
contiguous slice and the only collectives are the benchmark's barrier and the max-over-ranks of the
device time.  Kept free of CUDA so the N>1 logic is testable with the gloo backend on CPU.
"""
from __future__ import annotations


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


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Max of a python float over the process group (identity when not initialised)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, world: int, seconds_max: float) -> float:
    """Whole-job throughput: units all ranks processed / max-over-ranks time (weak scaling)."""
    return units_per_rank * world / seconds_max
