"""One resident-input pass of a BASELINE config at a reduced batch (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads
cfg, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
bb = sg.BatchBuilder(u_dtype=np.float32)
for kw in workloads.CONFIGS[cfg](n=n):
    bb.add_soundgen(**kw)
bt = sg.Batch()
bt.upload(bb.build())
for _ in range(reps):
    info = bt.run()
print({k: round(v, 3) for k, v in zip(['h2d', 'control', 'ampl', 'synth', 'compose', 'noise', 'assemble', 'envelope', 'filter', 'finalize', 'd2h', 'total'], info.ms)})
