"""world_size-2 CPU (gloo) test of the multi-GPU plumbing: per-rank workloads differ, shards tile
the batch, timings reduce with MAX and audio seconds with SUM -- no collective on the data path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soundgen_beta_b200 import sharding, workloads


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    calls = workloads.config3(n=2, seed=sharding.shard_seed(3, rank))
    fingerprint = float(calls[0]['rolloff'])
    dt, dte, audio = sharding.aggregate(dist, 1.0 + rank, 2.0 - rank, 10.0 * (rank + 1))
    q.put((rank, fingerprint, dt, dte, audio, sharding.shard_range(11, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] != res[1][1]                       # different synthetic batches per rank
    for r in res:
        assert (r[2], r[3], r[4]) == (2.0, 2.0, 30.0)   # MAX, MAX, SUM on every rank
    assert res[0][5] == (0, 6) and res[1][5] == (6, 11)


def test_shard_range_covers_everything():
    for n in (1, 7, 8192):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
