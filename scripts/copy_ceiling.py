"""Bare pinned-copy ceiling of a multi-GPU box: every rank copies 1 GiB device -> host (and host -> device)
from / to its own pinned buffer at the same time, nothing else running.  Launch like bench.py:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/copy_ceiling.py
Prints the aggregate GB/s: the ceiling the end-to-end fetch of bench.py can be compared with."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 1 << 28     # floats = 1 GiB
dev = torch.empty(n, dtype=torch.float32, device='cuda')
host = torch.empty(n, dtype=torch.float32).pin_memory()
res = {}
for name, (dst, src) in {'d2h': (host, dev), 'h2d': (dev, host)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = 8 * n * 4 * world / float(t.item()) / 1e9
if rank == 0:
    print(json.dumps({'n_gpus': world, 'aggregate_GBps': res, 'per_gpu_GBps': {k: v / world for k, v in res.items()},
                      'bytes_per_copy': n * 4, 'host_cores': os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
