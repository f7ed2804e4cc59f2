"""Multi-GPU plumbing of the benchmark / batch drivers: the path shards trivially (independent
sounds), so ranks never exchange samples; torch.distributed only carries timings."""
from __future__ import annotations


def shard_seed(config: int, rank: int) -> int:
    """Seed of the synthetic workload of `rank` (weak scaling: every rank gets its own batch)."""
    return 20260000 + config + 1000 * rank


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [begin, end) share of `n_items` calls for `rank` (strong-scaling helper)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def aggregate(dist, dt: float, dt_e2e: float, audio_s: float, device=None):
    """max over ranks of the two timings, sum over ranks of the audio seconds."""
    if dist is None:
        return dt, dt_e2e, audio_s
    import torch
    t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=device)
    a = torch.tensor([audio_s], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), float(a[0])
