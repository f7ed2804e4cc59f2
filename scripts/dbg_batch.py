"""Debug helper: run slices [lo, hi) of a config's batch, printing stage times (flushes before each run).
   python scripts/dbg_batch.py CFG TOTAL lo:hi [lo:hi ...]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads, _abi

cfg, total = int(sys.argv[1]), int(sys.argv[2])
calls = workloads.CONFIGS[cfg](n=total)
for spec in sys.argv[3:]:
    lo, hi = (int(v) for v in spec.split(':'))
    bb = sg.BatchBuilder(u_dtype=np.float32)
    for kw in calls[lo:hi]:
        bb.add_soundgen(**kw)
    bt = sg.Batch()
    bt.upload(bb.build())
    print('run', spec, '...', flush=True)
    t = time.time()
    info = bt.run()
    print('  ok %.2fs' % (time.time() - t), {n: round(v, 3) for n, v in zip(_abi.T_NAMES, info.ms)}, flush=True)
    bt.close()
