"""Replays bench.py's pipelined section many times, recording per-call lengths of every step."""
import os, sys, threading, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads, _abi, sharding
import bench

L = _abi.load()
n, npipe, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
with_main = len(sys.argv) > 4 and 'main' in sys.argv[4]
with_sampler = len(sys.argv) > 4 and 'smi' in sys.argv[4]
calls = workloads.CONFIGS[3](n=n, seed=sharding.shard_seed(3, 0))

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

if with_main:
    bt = sg.Batch(); bt.upload(build(0, n))
    for _ in range(3): bt.run()
    lens = bt.lengths(); out = np.zeros(int(lens.sum()), dtype=np.float32); L.sgb_pin(out.ctypes.data, out.nbytes)
    bt.fetch(np.float32, out=out)
    for _ in range(3): bt.run()
if with_sampler:
    smp = bench.ClockSampler(0); smp.start(); time.sleep(0.3)
subs = [build(*sharding.shard_range(n, i, npipe)) for i in range(npipe)]
batches = [sg.Batch() for _ in subs]
hist = [[] for _ in subs]
csum = [[] for _ in subs]
outs = [None] * npipe

def work(i):
    bt = batches[i]
    bt.upload(subs[i]); bt.run()
    lens = bt.lengths()
    hist[i].append(lens.copy())
    cs = np.zeros(9, dtype=np.uint64); L.sgb_batch_checksums(bt.h, cs.ctypes.data, 9); csum[i].append(cs)
    if outs[i] is None:
        outs[i] = np.zeros(int(lens.sum()) + 100000, dtype=np.float32)
        L.sgb_pin(outs[i].ctypes.data, outs[i].nbytes)
    bt.fetch(np.float32, out=outs[i])

for r in range(reps):
    th = [threading.Thread(target=work, args=(i,)) for i in range(npipe)]
    for t in th: t.start()
    for t in th: t.join()
bad = 0
for i in range(npipe):
    for r in range(1, reps):
        d = np.nonzero(hist[i][r] != hist[i][0])[0]
        if d.size:
            bad += 1
            print('sub', i, 'rep', r, 'calls', d[:6].tolist(), 'len', hist[i][r][d[:6]].tolist(), 'vs', hist[i][0][d[:6]].tolist())
            c = int(d[0])
            art = batches[i].artefacts(c)
            print('   last-run artefacts of call', c, {k: (v.tolist() if hasattr(v, 'tolist') and np.size(v) < 40 else None) for k, v in art.items() if k in ('zc', 'epochs', 'status', 'z_used')})
names = ['ctrl', 'pieces', 'tiles', 'amp32', 'amp64', 'wave', 'raw', 'sound', 'out']
for i in range(npipe):
    for r in range(1, reps):
        d = [names[k] for k in range(9) if csum[i][r][k] != csum[i][0][k]]
        if d:
            print('checksum diff sub', i, 'rep', r, d)
print('bad:', bad)
