"""Host-side logic of the Python mirror: argument munging, batch layout, workloads."""
import numpy as np
import pytest

import soundgen_beta_b200 as sg
from soundgen_beta_b200 import host, workloads
from oracle.rprims import r_spline, r_seq_len_out
from oracle import soundgen_oracle as so


def test_host_primitives_match_oracle():
    r = np.random.default_rng(0)
    for n in (2, 3, 4, 12):
        x = np.sort(r.uniform(0, 1, n)); x[0], x[-1] = 0, 1
        y = r.uniform(-5, 5, n)
        assert np.array_equal(host.spline(y, 57, x=x), r_spline(y, 57, x=x))
    assert np.array_equal(host.seq_len(1, 7, 13), r_seq_len_out(1, 7, 13))
    an = (np.linspace(0, 1, 12), r.uniform(80, 300, 12))
    a = host.smooth_contour(an, 1750, thisIsPitch=True, method='spline', valueFloor=50, valueCeiling=3500)
    b = so.getSmoothContour(an, length=1750, thisIsPitch=True, method='spline', valueFloor=50, valueCeiling=3500)
    assert np.array_equal(a, b)
    with pytest.raises(NotImplementedError):
        host.smooth_contour((np.linspace(0, 1, 4), np.ones(4)), 100)


def test_builder_layout():
    bb = sg.BatchBuilder()
    c = bb.add_soundgen(sylLen=200, nSyl=3, pauseLen=100, repeatBout=2, pitchAnchors=[100, 150], temperature=0,
                        noiseAnchors=((-50., 250.), (-30., -10.)), u=[np.zeros(bb.noise_uniform_count(4800, 800))] * 6)
    assert c == 0 and len(bb.bouts) == 2 and len(bb.syls) == 6 and len(bb.noises) == 6
    b0, b1 = bb.bouts
    assert (b0.syl_begin, b0.syl_end, b1.syl_begin, b1.syl_end) == (0, 3, 3, 6)
    assert b0.lead_silence == 1600 and b0.tail_silence == 0 and b1.lead_silence == 1600 and b1.tail_silence == 1600
    assert bb.syls[0].pause_after == 1616 and bb.syls[2].pause_after == 0   # gap = pauseLen + 1 ms (utilities_soundgen.R:546-549)
    assert bb.noises[0].len == 4800 and bb.noises[0].insertion == 1 - 800   # 50 ms pre-aspiration (soundgen.R:522-528)
    d = bb.build()
    assert d.n_calls == 1 and d.n_syllables == 6 and d.n_pitch == 6 * 700


def test_range_check_and_unsupported():
    bb = sg.BatchBuilder()
    w = []
    bb.add_soundgen(samplingRate=48000, temperature=0, pitchAnchors=[100, 150], warn=w)
    assert bb.syls[0].samplingRate == 16000 and any('samplingRate' in m for m in w)
    bb.add_soundgen(samplingRate=48000, temperature=0, pitchAnchors=[100, 150], invalidArgAction='ignore')
    assert bb.syls[1].samplingRate == 48000
    with pytest.raises(ValueError):
        bb.add_soundgen(sylLen=10, temperature=0, invalidArgAction='abort')
    with pytest.raises(NotImplementedError):
        bb.add_soundgen(temperature=0.1)


def test_workloads_are_seeded():
    a, b = workloads.config3(n=3), workloads.config3(n=3)
    assert all(np.array_equal(x['z'][0], y['z'][0]) and x['rolloff'] == y['rolloff'] for x, y in zip(a, b))
    c2 = workloads.config2(n=2)
    assert c2[0]['u'][0].size == workloads.noise_uniform_count(
        int(np.rint((c2[0]['noiseAnchors'][0].max() - c2[0]['noiseAnchors'][0].min()) * 22.05)), 1102)
