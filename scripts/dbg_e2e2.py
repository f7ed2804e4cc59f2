"""e2e / resident throughput of PipelinedBatches.run_steps for several (sub-batches, runners)."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(150, exit=True)
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads, _abi, sharding
L = _abi.load()
n = int(sys.argv[1])
calls = workloads.CONFIGS[3](n=n)
def build(lo, hi):
    bb = sg.BatchBuilder(u_dtype=np.float32)
    for kw in calls[lo:hi]:
        bb.add_soundgen(**kw)
    d = bb.build()
    sg.pin_desc(d)
    return d
for spec in sys.argv[2:]:
    npipe, runners = (int(v) for v in spec.split(':'))
    pipe = sg.PipelinedBatches([build(*sharding.shard_range(n, i, npipe)) for i in range(npipe)], runners=runners)
    pipe.run_steps(2)
    t0 = time.perf_counter(); pipe.run_steps(3); te = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter(); pipe.run_steps(3, transfer=False); tr = (time.perf_counter() - t0) / 3
    print('sub-batches %d runners %d: e2e %.1f ms/step, resident %.1f ms/step' % (npipe, runners, te * 1e3, tr * 1e3), flush=True)
    pipe.close()
    for d in pipe.descs: sg.pin_desc(d, False)
