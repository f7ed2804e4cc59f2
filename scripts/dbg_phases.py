"""Host-side wall time of the three phases of a front-end run (cfg4 presets): run_begin / resolve / run_finish."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads, _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
t = time.perf_counter()
bb = sg.BatchBuilder(u_dtype=np.float32)
for kw in workloads.config4(n=n):
    bb.add_soundgen(**kw)
d = bb.build()
print('front-end build %.1f ms' % ((time.perf_counter() - t) * 1e3), 'threads', os.cpu_count())
bt = sg.Batch(); bt.upload(d)
L = _abi.load()
for rep in range(3):
    t0 = time.perf_counter(); L.sgb_batch_run_begin(bt.h)
    t1 = time.perf_counter(); d._fe.resolve(bt)
    t2 = time.perf_counter(); info = _abi.RunInfo(); L.sgb_batch_run_finish(bt.h, C.byref(info))
    t3 = time.perf_counter(); d._fe.round_end(bt)
    t4 = time.perf_counter()
    print('begin %.1f  resolve %.1f  finish %.1f  round_end %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
