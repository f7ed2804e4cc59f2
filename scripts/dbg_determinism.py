"""Run the same batch several times (one handle, and several handles on threads) and compare results."""
import os, sys, threading
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import soundgen_beta_b200 as sg
from soundgen_beta_b200 import workloads

cfg, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nthreads = int(sys.argv[4]) if len(sys.argv) > 4 else 1
pin = len(sys.argv) > 5 and 'pin' in sys.argv[5]
big = len(sys.argv) > 5 and 'big' in sys.argv[5]
from soundgen_beta_b200 import _abi
L = _abi.load()
calls = workloads.CONFIGS[cfg](n=n)


def build(lo, hi):
    bb = sg.BatchBuilder(u_dtype=np.float32)
    for kw in calls[lo:hi]:
        bb.add_soundgen(**kw)
    d = bb.build()
    if pin:
        for k in ('pitch', 'anchors', 'formants', 'z', 'u', 'pre'):
            a = d._keep[k]
            if a.size:
                L.sgb_pin(a.ctypes.data, a.nbytes)
    return d


def worker(tid, lo, hi, res):
    bt = sg.Batch()
    d = build(lo, hi)
    ref = None
    for r in range(reps):
        bt.upload(d)
        bt.run()
        lens = bt.lengths()
        out = bt.fetch(np.float32)
        if ref is None:
            ref = (lens.copy(), [o.copy() for o in out])
            continue
        bad_len = np.nonzero(lens != ref[0])[0]
        bad_val = [i for i in range(len(out)) if i not in set(bad_len) and not np.array_equal(out[i], ref[1][i])]
        if len(bad_len) or bad_val:
            res.append((tid, r, bad_len[:5].tolist(), [(int(lens[i]), int(ref[0][i])) for i in bad_len[:5]], bad_val[:5],
                        [float(np.max(np.abs(out[i] - ref[1][i]))) for i in bad_val[:5]]))
    bt.close()


res = []
th = []
if big:
    bt0 = sg.Batch(); bt0.upload(build(0, n)); bt0.run(); bt0.run()
per = n // nthreads
for t in range(nthreads):
    th.append(threading.Thread(target=worker, args=(t, t * per, (t + 1) * per, res)))
for t in th: t.start()
for t in th: t.join()
print('mismatches:', len(res))
for r in res[:20]:
    print(r)
