"""Run-to-run reproducibility of the whole path with several batches in flight on their own
streams (the way bench.py's end-to-end measurement and a multi-threaded R front end drive the
library): identical inputs must give bit-identical intermediates and outputs every time."""
import threading

import numpy as np
import pytest

import soundgen_beta_b200 as sg
from soundgen_beta_b200 import _abi, sharding, workloads

pytestmark = pytest.mark.gpu
NAMES = ['ctrl', 'pieces', 'tiles', 'amp32', 'amp64', 'wave', 'raw', 'sound', 'out']


@pytest.mark.parametrize('cfg,n,npipe,reps', [(3, 2048, 8, 5), (1, 512, 4, 4), (4, 1056, 8, 5)])
def test_pipelined_runs_are_bit_reproducible(cfg, n, npipe, reps):
    L = _abi.load()
    calls = workloads.CONFIGS[cfg](n=n)
    descs = []
    for i in range(npipe):
        lo, hi = sharding.shard_range(n, i, npipe)
        bb = sg.BatchBuilder(u_dtype=np.float32)
        for kw in calls[lo:hi]:   # cfg4 keeps its temperature: the draws that wait for device results (formant
            bb.add_soundgen(**kw)  # tracks, sgb_frontend_resolve) restart from the same stream position on every run
        descs.append(bb.build())
    batches = [sg.Batch() for _ in descs]
    sums = [[] for _ in descs]
    lens = [[] for _ in descs]
    errs = []

    def work(i):
        try:
            bt = batches[i]
            bt.upload(descs[i])
            bt.run()
            cs = np.zeros(9, dtype=np.uint64)
            assert L.sgb_batch_checksums(bt.h, cs.ctypes.data, 9) == 0
            sums[i].append(cs)
            lens[i].append(bt.lengths().copy())
            bt.fetch(np.float32)
        except Exception as e:   # noqa: BLE001 - surfaced below
            errs.append(e)

    for _ in range(reps):
        th = [threading.Thread(target=work, args=(i,)) for i in range(npipe)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs[0]
    for i in range(npipe):
        for r in range(1, reps):
            assert np.array_equal(lens[i][r], lens[i][0]), (i, r)
            diff = [NAMES[k] for k in range(9) if sums[i][r][k] != sums[i][0][k]]
            assert not diff, 'sub-batch %d, repetition %d: %s differ' % (i, r, diff)
    for b in batches:
        b.close()


@pytest.mark.parametrize('cfg,n', [(3, 96), (4, 66)])
def test_pipeline_from_argument_lists_equals_prebuilt(cfg, n):
    """Every step of PipelinedBatches(sources=...) runs the library front-end on the argument structs
    (sgb_frontend_add_many on worker threads): same waveforms as the prebuilt descriptions, step after step."""
    calls = workloads.CONFIGS[cfg](n=n)
    npipe = 3
    descs, srcs = [], []
    for i in range(npipe):
        lo, hi = sharding.shard_range(n, i, npipe)
        bb = sg.BatchBuilder(u_dtype=np.float32)
        for kw in calls[lo:hi]:
            bb.add_soundgen(**kw)
        descs.append(bb.build())
        srcs.append(sg.ArgArray(calls[lo:hi], np.float32))
    ref = sg.PipelinedBatches(descs, runners=2)
    want = [np.array(w, copy=True) for w in ref.step(np.float32)]
    ref.close()
    pipe = sg.PipelinedBatches(sources=srcs, runners=2, fe_threads=2)
    for _ in range(3):
        got = pipe.step(np.float32)
        assert len(got) == len(want) == n
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    pipe.close()
