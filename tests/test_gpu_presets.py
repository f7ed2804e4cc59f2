"""BASELINE config index 4: every bundled preset through the CUDA path against the oracle
(temperature 0, spline contours: the stochastic host stage and loess need R, see workloads.config4)."""
import numpy as np
import pytest

import soundgen_beta_b200 as sg
from oracle import soundgen_oracle as so
from oracle.soundgen_call import soundgen as osg
from soundgen_beta_b200 import presets, workloads

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_all_presets_match_the_oracle():
    calls = workloads.config4(n=33)
    bb = sg.BatchBuilder()
    logs = []
    for kw in calls:
        kw = dict(kw)
        z, u = workloads.streams(kw.pop('seed'))
        zl, ul = [], []
        s0 = len(bb.syls)
        bb.add_soundgen(z=lambda n, z=z, zl=zl: (zl.append(z(n)) or zl[-1]),
                        u=lambda n, u=u, ul=ul: (ul.append(u(n)) or ul[-1]), **kw)
        voiced = [s for s in range(s0, len(bb.syls)) if bb.syls[s].kind == 1]
        logs.append((zl, ul, voiced))
    bt = sg.Batch()
    bt.upload(bb.build())
    bt.run()
    assert np.all(bt.status() == 0)
    outs = bt.fetch(np.float64)
    worst = 0.0
    for (spk, name, _), kw, (zl, ul, voiced), y in zip(presets.load(), calls, logs, outs):
        kw = dict(kw)
        kw.pop('seed')
        used = [bt.artefacts(s)['z_used'] for s in voiced]
        zcat = np.concatenate([z[:n] for z, n in zip(zl, used)]) if zl else None
        ucat = np.concatenate(ul) if ul else None
        ref = osg(rng=so.RStream(z=zcat, u=ucat), **kw)
        assert y.size == ref.size, (spk, name, y.size, ref.size)
        err = float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
        worst = max(worst, err)
        assert err < TOL, (spk, name, err)
    print('worst preset error %.2e of peak' % worst)


def test_ragged_batch_of_unrelated_calls():
    """One batch mixing sampling rates, window lengths, voiced / unvoiced / multi-bout calls: every call must
    come out as if it had been run alone (the oracle runs them one by one)."""
    ps = {(s, n): kw for s, n, kw in presets.load()}
    picks = [('Cat', 'Hiss'), ('Misc', 'Seagull'), ('M1', 'Sigh'), ('Cat', 'Purr'), ('Misc', 'Duck')]
    calls = [dict(ps[p], temperature=0, contour_method='spline', seed=900 + i) for i, p in enumerate(picks)]
    extra = workloads.config1(n=2) + workloads.config3(n=1) + workloads.config0()
    bb = sg.BatchBuilder()
    logs = []
    for kw in calls:
        kw = dict(kw)
        z, u = workloads.streams(kw.pop('seed'))
        zl, ul = [], []
        s0 = len(bb.syls)
        bb.add_soundgen(z=lambda n, z=z, zl=zl: (zl.append(z(n)) or zl[-1]),
                        u=lambda n, u=u, ul=ul: (ul.append(u(n)) or ul[-1]), **kw)
        logs.append((zl, ul, [s for s in range(s0, len(bb.syls)) if bb.syls[s].kind == 1]))
    for kw in extra:
        bb.add_soundgen(**kw)
    bt = sg.Batch()
    bt.upload(bb.build())
    bt.run()
    assert np.all(bt.status() == 0)
    outs = bt.fetch(np.float64)
    for kw, (zl, ul, voiced), y in zip(calls, logs, outs):
        kw = dict(kw)
        kw.pop('seed')
        used = [bt.artefacts(s)['z_used'] for s in voiced]
        ref = osg(rng=so.RStream(z=np.concatenate([z[:n] for z, n in zip(zl, used)]) if zl else None,
                                 u=np.concatenate(ul) if ul else None), **kw)
        assert y.size == ref.size and float(np.max(np.abs(y - ref)) / np.max(np.abs(ref))) < TOL
    for kw, y in zip(extra, outs[len(calls):]):
        kw = dict(kw)
        z, u = kw.pop('z', None), kw.pop('u', None)
        ref = osg(rng=so.RStream(z=np.concatenate(z) if z else None, u=np.concatenate(u) if u else None), **kw)
        assert y.size == ref.size and float(np.max(np.abs(y - ref)) / np.max(np.abs(ref))) < TOL
