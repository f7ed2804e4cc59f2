"""Host-side handling of the bundled presets (BASELINE config index 4): the data table is complete and
every preset goes through the soundgen() argument munging without a GPU."""
import numpy as np

import soundgen_beta_b200 as sg
from soundgen_beta_b200 import presets, workloads


def test_preset_table():
    ps = presets.load()
    assert len(ps) == 33                                   # R/presets.R:156-410
    assert [s for s, _, _ in ps].count('Cat') == 9 and ps[0][:2] == ('M1', 'Vowel1')
    kw = presets.preset('Misc', 'Seagull')
    assert kw['samplingRate'] == 24000 and kw['nSyl'] == 8
    assert presets.preset('Cat', 'Hiss')['pitchAnchors'] is None          # NULL: unvoiced
    assert presets.preset('Misc', 'Elephant')['formants'] is None         # NA: schwa from vocalTract


def test_config4_builds_on_the_host():
    calls = workloads.config4(n=66)
    bb = sg.BatchBuilder(u_dtype=np.float32)
    for kw in calls:
        kw = dict(kw)
        z, u = workloads.streams(kw.pop('seed'), np.float32)
        bb.add_soundgen(z=z, u=u, **kw)
    assert len(bb.calls) == 66 and len(bb.syls) == 2 * 63 and len(bb.noises) == 2 * 48
    d = bb.build()
    assert d.n_calls == 66 and d.n_u == sum(a.size for a in bb.u)
