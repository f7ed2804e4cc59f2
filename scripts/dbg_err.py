"""max |err| / peak of the CUDA path against the oracle for the first n calls of a config."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads
from oracle import soundgen_oracle as so
from oracle.soundgen_call import soundgen as osg

cfg, n = int(sys.argv[1]), int(sys.argv[2])
calls = workloads.CONFIGS[cfg](n=n) if cfg else workloads.config0()
outs, st = sg.soundgen_batch(calls, out_dtype=np.float64)
errs = []
for kw, y in zip(calls, outs):
    kw = dict(kw); z, u = kw.pop('z', None), kw.pop('u', None)
    ref = osg(rng=so.RStream(z=np.concatenate(z) if z else None, u=np.concatenate(u) if u else None), **kw)
    errs.append(float(np.max(np.abs(y - ref)) / np.max(np.abs(ref))) if y.size == ref.size else float('nan'))
print('cfg', cfg, 'mode', os.environ.get('SGB_SYNTH_MODE', 'auto'), 'max rel err per call:', ' '.join('%.2e' % e for e in errs))
