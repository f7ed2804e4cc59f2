"""Wall-clock anatomy of the pipelined end-to-end path: per sub-batch upload / run / fetch times."""
import os, sys, threading, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads, _abi, sharding

L = _abi.load()
n, npipe, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
W = int(sys.argv[4]) if len(sys.argv) > 4 else npipe
calls = workloads.CONFIGS[3](n=n)
print('cpus', os.cpu_count(), flush=True)

def build(lo, hi):
    bb = sg.BatchBuilder(u_dtype=np.float32)
    for kw in calls[lo:hi]:
        bb.add_soundgen(**kw)
    d = bb.build()
    for k in ('pitch', 'anchors', 'formants', 'z', 'u', 'pre'):
        a = d._keep[k]
        if a.size:
            L.sgb_pin(a.ctypes.data, a.nbytes)
    return d

subs = [build(*sharding.shard_range(n, i, npipe)) for i in range(npipe)]
batches = [sg.Batch() for _ in subs]
outs = [None] * npipe
log = []

def work(i, r, t0):
    bt = batches[i]
    a = time.perf_counter(); bt.upload(subs[i])
    b = time.perf_counter(); info = bt.run()
    c = time.perf_counter()
    if outs[i] is None:
        outs[i] = np.zeros(int(bt.lengths().sum()), dtype=np.float32); L.sgb_pin(outs[i].ctypes.data, outs[i].nbytes)
    c2 = time.perf_counter(); bt.fetch(np.float32, out=outs[i])
    d = time.perf_counter()
    log.append((r, i, a - t0, b - a, c - b, d - c2, sum(info.ms[1:11]), outs[i].nbytes / (d - c2) / 1e9))

def worker(w, t0, nsteps):
    for r in range(nsteps):
        for i in range(w, npipe, W):
            work(i, r, t0)

for i in range(npipe):   # warm-up: size pools, pin outputs
    work(i, -1, time.perf_counter())
for nsteps in (1, reps):
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(w, t0, nsteps)) for w in range(W)]
    for t in th: t.start()
    for t in th: t.join()
    print('npipe', npipe, 'workers', W, 'steps', nsteps, 'wall per step %.1f ms' % ((time.perf_counter() - t0) * 1e3 / nsteps), flush=True)
for e in (log[-npipe:] if len(sys.argv) > 5 else []):
    print('rep %d sub %d start %.1f upload %.1f run %.1f (gpu stages %.1f) fetch %.1f ms (%.1f GB/s)' % (e[0], e[1], e[2] * 1e3, e[3] * 1e3, e[4] * 1e3, e[6], e[5] * 1e3, e[7]))
