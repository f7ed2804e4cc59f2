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
    """Every preset, with its own temperature under set.seed(i), goes through the host front-end."""
    calls = workloads.config4(n=66)
    bb = sg.BatchBuilder(u_dtype=np.float32)
    for kw in calls:
        bb.add_soundgen(**kw)
    d = bb.build()
    assert d.n_calls == 66 and d.n_syllables >= 66 and d.n_noises > 40
    assert bb.h2d_bytes() > 4 * d.n_u
