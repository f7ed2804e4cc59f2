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


def aggregate(dist, dt: float, dt_e2e: float, audio_s: float, device=None, extra=()):
    """max over ranks of the timings (dt, dt_e2e, *extra), sum over ranks of the audio seconds."""
    if dist is None:
        return (dt, dt_e2e, audio_s) + tuple(extra) if extra else (dt, dt_e2e, audio_s)
    import torch
    t = torch.tensor([dt, dt_e2e] + list(extra), dtype=torch.float64, device=device)
    a = torch.tensor([audio_s], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    out = (float(t[0]), float(t[1]), float(a[0]))
    return out + tuple(float(v) for v in t[2:]) if extra else out


def bind_to_gpu_numa(L, device: int) -> list:
    """Pins the calling process (and the threads it starts later) to the CPUs NVML reports as local to CUDA
    device `device`, so that a rank's pinned buffers and its upload / fetch threads live on the GPU's NUMA
    node.  Best effort: returns the CPU list, [] when NVML or the affinity call is unavailable."""
    import ctypes
    import os
    try:
        import pynvml
        buf = ctypes.create_string_buffer(32)
        if L.sgb_device_pci_bus_id(device, buf, 32) != 0:
            return []
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(buf.value)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []
