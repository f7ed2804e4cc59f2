"""Resident throughput of cfg3 when the batch runs as sub-batches on several streams (no transfers in the loop)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads, sharding
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
calls = workloads.CONFIGS[cfg](n=n)
srs = np.array([float(kw.get('samplingRate', 16000)) for kw in calls])
for npipe, runners in ((1, 1), (2, 2), (4, 2), (4, 4), (8, 3), (8, 4), (16, 4)):
    descs = []
    for i in range(npipe):
        lo, hi = sharding.shard_range(n, i, npipe)
        bb = sg.BatchBuilder(u_dtype=np.float32)
        for kw in calls[lo:hi]:
            bb.add_soundgen(**kw)
        descs.append(bb.build())
    pipe = sg.PipelinedBatches(descs, runners=runners)
    pipe.run_steps(2, transfer=True)          # upload once, size the pools
    audio = sum(float(np.sum(b.lengths() / srs[sharding.shard_range(n, i, npipe)[0]:sharding.shard_range(n, i, npipe)[1]])) for i, b in enumerate(pipe.batches))
    steps = 6
    t = time.perf_counter()
    pipe.run_steps(steps, transfer=False)
    dt = time.perf_counter() - t
    print('sub-batches %2d runners %d: %.1f ms per step, %.0f audio-s/s' % (npipe, runners, dt / steps * 1e3, audio * steps / dt), flush=True)
    pipe.close()
