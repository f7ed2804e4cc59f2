"""The oracle against the known answers of SURVEY.md Appendix B (tests/golden/appendix_b.json)
and structural invariants.  The reference ships no golden vectors (parity unpinned)."""
import json
import os

import numpy as np

from oracle import soundgen_oracle as so
from oracle.rprims import fmm_coef, fmm_eval, r_approx, r_seq_by, r_seq_len_out, r_spline
from oracle.soundgen_call import soundgen

G = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'appendix_b.json')))


def test_glottal_cycles_and_upsample():
    gc = so.getGlottalCycles(r_seq_len_out(150, 200, 350), 3500)
    assert gc.tolist() == G['getGlottalCycles_150_200_350_3500']
    pu, gcu = so.upsample(np.array([100, 150, 130.]), 16000)
    assert gcu.tolist() == G['upsample_100_150_130_16000_gc'] and pu.size == 390


def test_small_utilities():
    s = G['clumper_in']
    assert so.clumper(s, 2).tolist() == G['clumper_2']
    assert so.clumper(s, 3).tolist() == G['clumper_3']
    assert so.addVectors(np.arange(1, 7), np.full(3, 100.), 5).tolist() == G['addVectors_5']
    assert so.addVectors(np.arange(1, 7), np.full(3, 100.), -4).tolist() == G['addVectors_m4']
    assert so.matchLengths([1, 2, 3], 5).tolist() == G['matchLengths_123_5']
    assert so.matchLengths([3, 4, 5, 6, 7], 3).tolist() == G['matchLengths_37_3']
    for k, (q1, q2) in G['noiseThresholds'].items():
        a, b = so.noise_thresholds(int(k))
        assert abs(a - q1) < 1e-4 and abs(b - q2) < 1e-4


def test_r_primitives():
    # FMM is exact on cubics for n >= 4 and is the interpolating parabola for n = 3
    x = np.array([0., 1., 2.5, 4., 6.])
    f = lambda t: 2 - t + 0.5 * t ** 2 - 0.1 * t ** 3
    b, c, d = fmm_coef(x, f(x))
    u = np.linspace(0, 6, 50)
    assert np.allclose(fmm_eval(x, f(x), b, c, d, u), f(u), atol=1e-12)
    x3 = np.array([0., 1., 3.])
    q = lambda t: 1 + 2 * t - 0.3 * t ** 2
    b, c, d = fmm_coef(x3, q(x3))
    assert np.allclose(fmm_eval(x3, q(x3), b, c, d, u[:25]), q(u[:25]), atol=1e-12)
    assert r_seq_len_out(0, 1, 5).tolist() == [0, .25, .5, .75, 1]
    assert r_seq_by(1, 10, 2.5).tolist() == [1, 3.5, 6, 8.5]
    assert np.allclose(r_approx([0, 10, 0], 5), [0, 5, 10, 5, 0])
    assert r_spline([1., 2.], 3).tolist() == [1, 1.5, 2]


def test_filter_geometry_and_gain():
    g = G['default_geometry']
    step = so.frame_starts(16000, g['wl'], 75)
    assert step.size == g['frames_16000']
    y = so.istft(so.stft_complex(np.ones(16000), 800, step), 800, 75)
    assert y.size == g['filtered_len_16000']
    assert abs(y[2000:14000].mean() - G['filter_gain_wl800']) < 5e-6
    step = so.frame_starts(48000, 2400, 75)
    y = so.istft(so.stft_complex(np.ones(48000), 2400, step), 2400, 75)
    assert y.size == 47400 and abs(y[5000:40000].mean() - G['filter_gain_wl2400']) < 2e-6
    step = so.frame_starts(44100, 1102, 75)   # non-integer hop: 275.5
    assert step.size == 157 and step[1] == 276.5
    assert so.istft(so.stft_complex(np.zeros(44100), 1102, step), 1102, 75).size == 44080


def test_default_call_known_answer():
    k = G['soundgen_1000ms_100_150']
    y, arts, _ = soundgen(sylLen=1000, pitchAnchors=[100, 150], temperature=0, addSilence=100,
                          want_artefacts=True)
    a = arts[0]
    assert a.gc[:6].tolist() == k['gc_head'] and a.gc.size == k['G']
    assert a.nHarmonics == k['nHarmonics'] and a.rows_kept == k['rows_kept']
    assert a.gc_upsampled[:6].tolist() == k['gc_upsampled_head']
    assert a.gc_upsampled[-2:].tolist() == k['gc_upsampled_tail']
    assert a.zc == [(1, k['zc2'])]
    assert y.size == k['final_len']
    assert abs(np.max(y) - 1.0) < 1e-12   # signed-max normalisation


def test_invariants_random():
    from cases import voiced_case
    for seed in range(6):
        pitch, z, anchors, pars = voiced_case(seed, 0.2)
        y, a = so.generateHarmonics(pitch, rng=so.RStream(z=z), amplAnchors=anchors, want_artefacts=True, **pars)
        assert np.all(np.diff(a.gc) >= 2) and np.all(np.diff(a.gc_upsampled) > 0)
        assert a.gc_upsampled[-1] == a.n_upsampled == a.integr.size
        assert np.all(np.diff(a.integr) > 0)
        assert a.epochs[0, 0] == 1 and a.epochs[-1, 1] == a.gc.size
        assert np.isfinite(y).all() and abs(np.max(y)) <= 1.0 + 0.5


def test_savewav_pcm16_known_answers():
    # seewave::savewav -> tuneR::normalize(unit = '16'): centred, max |x| -> level = min(1, max(x)), round half even
    x = np.array([0.0, 0.5, -0.5, 0.25])
    q = so.savewav_pcm16(x)                  # mean 0.0625; centred max |.| = 0.5625; level = 0.5
    assert q.tolist() == [int(np.rint(0.5 * (v - 0.0625) / 0.5625 * 32767)) for v in x]
    assert so.savewav_pcm16(np.array([2.0, -2.0])).tolist() == [32767, -32767]      # level capped at 1
    assert so.savewav_pcm16(np.zeros(4)).tolist() == [0, 0, 0, 0]                    # all.equal(m, 0): no scaling
