"""Parity of the CUDA path (through the C ABI) with the oracle, on a B200.
Tolerance (BASELINE.json north_star): max |err| <= 1e-4 of peak per waveform; integer artefacts
(glottal cycles, gc_upsampled, epochs, zero crossings, lengths) bit-exact."""
import numpy as np
import pytest

import soundgen_beta_b200 as sg
from cases import voiced_case
from oracle import soundgen_oracle as so
from oracle.soundgen_call import soundgen as osg
from soundgen_beta_b200 import workloads

pytestmark = pytest.mark.gpu
TOL = 1e-4
FM = [np.array([[0, 860, 30, 120.]]), np.array([[0, 1280, 40, 120.]]), np.array([[0, 2900, 25, 200.]])]
FM2 = [np.array([[0, 860, 30, 120.], [1, 500, 35, 100]]), np.array([[0, 1280, 40, 120.], [1, 2000, 20, 150]])]


def rel(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_rolloff_and_envelope():
    for kw in (dict(pitch_per_gc=[150, 800, 3000], rolloffOct=0), dict(pitch_per_gc=[150], rolloffParab=30, rolloffParabHarm=4),
               dict(pitch_per_gc=[150, 600], rolloffParab=-20, rolloffParabCeiling=2000),
               dict(pitch_per_gc=np.linspace(100, 400, 50), rolloff=np.linspace(-20, -6, 50), rolloffKHz=-3)):
        a, (b, _) = sg.getRolloff(**kw), so.getRolloff(**kw)
        assert a.shape == b.shape and rel(a, b) < 1e-12
    mo = (np.array([0., 1]), np.array([0, .8]))
    a = sg.getSpectralEnvelope(551, 40, formants=FM2, vocalTract=15.5, mouthAnchors=mo)
    b = so.getSpectralEnvelope(551, 40, formants=FM2, vocalTract=15.5, mouthAnchors=mo)
    assert rel(a, b) < 1e-10
    a = sg.getSpectralEnvelope(400, 1, formants=FM, vocalTract=15.5, mouthAnchors=[.5, .5])
    b = so.getSpectralEnvelope(400, 1, formants=FM, vocalTract=15.5, mouthAnchors=(np.array([0., 1]), np.array([.5, .5])))
    assert rel(a, b) < 1e-10
    assert rel(sg.getSpectralEnvelope(256, 3, formants=None, vocalTract=None), so.getSpectralEnvelope(256, 3)) < 1e-12


@pytest.mark.parametrize('T', [0.0, 0.3])
def test_generate_harmonics(T):
    for seed in range(10):
        pitch, z, anchors, pars = voiced_case(seed, T)
        ref, art = so.generateHarmonics(pitch, rng=so.RStream(z=z), amplAnchors=anchors, want_artefacts=True, **pars)
        y, a = sg.generateHarmonics(pitch, z=z, amplAnchors=anchors, want_artefacts=True, **pars)
        assert np.array_equal(a['gc'], art.gc) and np.array_equal(a['gc_upsampled'], art.gc_upsampled)
        assert np.array_equal(a['epochs'], art.epochs) and a['z_used'] == art.z_used
        assert [tuple(int(v or 0) for v in zz) for zz in art.zc] == [tuple(r) for r in a['zc'].tolist()]
        if art.jitter_idx is not None:
            assert np.array_equal(a['jitter_idx'], art.jitter_idx)
        if art.rw_bin is not None:
            assert np.array_equal(a['rw_bin'], art.rw_bin.astype(int))
        assert y.size == ref.size and rel(y, ref) < TOL


@pytest.mark.parametrize('f0a,f0b,sr', [(2400., 3300., 16000.), (900., 2600., 22050.), (55., 70., 48000.)])
def test_generate_harmonics_pitch_extremes(f0a, f0b, sr):
    # very high pitch: a warp's 128 samples span more than 16 glottal cycles (K1's direct, unstaged path and
    # the short staged blocks); very low pitch at 48 kHz: hundreds of rows, one cycle per warp
    P = 1400
    pitch = np.exp(np.linspace(np.log(f0a), np.log(f0b), P))
    z = np.random.default_rng(int(f0a)).standard_normal(20000)
    pars = dict(samplingRate=sr, pitchFloor=50, pitchCeiling=3500, nonlinBalance=100, jitterDep=1.0, jitterLen=5,
                shimmerDep=10, rolloff=-6, rolloffOct=-1, rolloffKHz=-2, subFreq=90, subDep=40, shortestEpoch=100,
                attackLen=10)
    ref, art = so.generateHarmonics(pitch, rng=so.RStream(z=z), want_artefacts=True, **pars)
    y, a = sg.generateHarmonics(pitch, z=z, want_artefacts=True, **pars)
    assert np.array_equal(a['gc_upsampled'], art.gc_upsampled) and np.array_equal(a['epochs'], art.epochs)
    assert y.size == ref.size and rel(y, ref) < TOL


@pytest.mark.parametrize('wl,n,moving,sr', [(800, 16000, False, 16000), (2204, 22050, True, 44100),
                                            (1102, 44100, False, 22050), (2400, 48000, True, 48000),
                                            (160, 5000, False, 16000), (1200, 30000, True, 24000),
                                            (4410, 30000, False, 44100)])
def test_filter(wl, n, moving, sr):
    x = np.random.default_rng(wl).standard_normal(n)
    nc = so.frame_starts(n, wl, 75).size
    env = so.getSpectralEnvelope(wl // 2, nc if moving else 1, formants=FM2 if moving else FM, vocalTract=15.5,
                                 samplingRate=sr)
    assert rel(sg.filter_sound(x, env, wl), so.filter_sound(x, env, wl)) < TOL


def test_filter_short_sound_clamps_window():
    x = np.random.default_rng(1).standard_normal(1204)     # wl -> floor(1204/2) = 602 = 2 * 7 * 43
    env = so.getSpectralEnvelope(301, 1, formants=FM, vocalTract=15.5)
    assert rel(sg.filter_sound(x, env, 800), so.filter_sound(x, env, 602)) < TOL


def test_filter_roundtrip_property():
    # unit envelope: the chain is a (nearly flat) gain, so after /max both sides agree with the input shape
    x = np.sin(np.arange(48000) * 0.05) + 0.3
    y = sg.filter_sound(x, np.ones(1200), 2400)
    ref = x[:y.size] / np.max(x[2400:y.size - 2400])
    assert np.max(np.abs(y[2400:-2400] - ref[2400:-2400])) < 2e-3   # window ripple ~3e-5, edges excluded


def test_generate_noise():
    for wl, n, ro in ((800, 8000, -6), (1102, 30000, -12), (2400, 5000, 0)):
        cnt = workloads.noise_uniform_count(n, wl)
        u = np.random.default_rng(wl).random(cnt)
        an = (np.array([0., 500]), np.array([-20., 10]))
        ref = so.generateNoise(n, an, rolloffNoise=ro, attackLen=10, windowLength_points=wl, rng=so.RStream(u=u))
        assert rel(sg.generateNoise(n, an, rolloffNoise=ro, attackLen=10, windowLength_points=wl, u=u), ref) < TOL


def test_generate_noise_with_a_literal_filter_matrix():
    # generateNoise(filterNoise = <matrix>) as soundgen() calls it (R/soundgen.R:668-690)
    wl, n = 800, 9000
    u = np.random.default_rng(11).random(workloads.noise_uniform_count(n, wl))
    an = (np.array([0., 400]), np.array([-10., 5]))
    for k in (1, 37):
        flt = so.getSpectralEnvelope(wl // 2, k, formants=FM2 if k > 1 else FM, vocalTract=15.5)
        ref = so.generateNoise(n, an, rolloffNoise=-8, attackLen=20, windowLength_points=wl, filterNoise=flt,
                               rng=so.RStream(u=u))
        y = sg.generateNoise(n, an, rolloffNoise=-8, attackLen=20, windowLength_points=wl, filterNoise=flt, u=u)
        assert rel(y, ref) < TOL


def test_pcm16_fetch_matches_savewav():
    # savePath branch (R/soundgen.R:855-857): seewave::savewav -> tuneR::normalize(unit = '16').  The device
    # quantises its FP32 waveform (error ~2e-6 of peak = 0.07 LSB), so a sample may land one step away from
    # the oracle's; one LSB is 3e-5 of peak, inside the 1e-4 bar.
    calls = workloads.config1(n=3) + workloads.config0()
    bb = sg.BatchBuilder()
    for kw in calls:
        bb.add_soundgen(**kw)
    bt = sg.Batch()
    bt.upload(bb.build())
    bt.run()
    pcm = bt.fetch(np.int16)
    for kw, q in zip(calls, pcm):
        ref = so.savewav_pcm16(_oracle_call(kw))
        assert q.dtype == np.int16 and q.size == ref.size
        d = np.abs(q.astype(np.int64) - ref)
        assert d.max() <= 1 and np.mean(d > 0) < 0.2
        assert np.max(np.abs(ref)) >= 32700          # the scaling really happened


def _oracle_call(kw):
    kw = dict(kw)
    z, u = kw.pop('z', None), kw.pop('u', None)
    if 'seed' in kw:                        # R's own stream after set.seed()
        from oracle.rrng import RRng
        rng = RRng(kw.pop('seed'))
    else:
        rng = so.RStream(z=np.concatenate(z) if z else None, u=np.concatenate(u) if u else None)
    return osg(rng=rng, **kw)


@pytest.mark.parametrize('cfg,n', [(0, 8), (1, 6), (2, 4), (3, 4)])
def test_configs_batched(cfg, n):
    calls = workloads.CONFIGS[cfg](n=n)      # cfg0: set.seed(1..8); soundgen(sylLen = 1000), the literal default call
    outs, st = sg.soundgen_batch(calls, out_dtype=np.float64)
    assert np.all(st == 0)
    for kw, y in zip(calls, outs):
        ref = _oracle_call(kw)
        assert y.size == ref.size and rel(y, ref) < TOL


def test_multisyllable_bouts_noise_am():
    bb = sg.BatchBuilder()
    cnt = bb.noise_uniform_count(int(round(350 * 16)), 800)
    r = np.random.default_rng(9)
    kw = dict(sylLen=200, nSyl=3, pauseLen=100, repeatBout=2, pitchAnchors=[180, 120], temperature=0,
              noiseAnchors=((-50., 300.), (-30., -5.)), amDep=40, amFreq=25, amShape=-0.3,
              pitchAnchorsGlobal=[0, 4, -2], amplAnchorsGlobal=[100, 120], nonlinBalance=100, jitterDep=1.5,
              shimmerDep=10, subFreq=80, subDep=60)
    us = [r.random(cnt) for _ in range(6)]
    zs = [r.standard_normal(600) for _ in range(6)]
    # the oracle consumes ONE stream per call: lay the per-syllable streams out in its order
    y, bt = sg.soundgen(z=zs, u=us, return_batch=True, **kw)
    used = [bt.artefacts(s)['z_used'] for s in range(6)]
    zcat = np.concatenate([z[:n] for z, n in zip(zs, used)])
    ref = osg(rng=so.RStream(z=zcat, u=np.concatenate(us)), **kw)
    assert y.size == ref.size and rel(y, ref) < TOL
    # post-filter noise (formantsNoise given) and moving mouth
    kw2 = dict(sylLen=300, pitchAnchors=[150, 220], temperature=0, noiseAnchors=((0., 400.), (-10., 0.)),
               formantsNoise=FM2, mouthAnchors=[0.2, 0.9], vocalTract=14)
    u2 = [r.random(bb.noise_uniform_count(int(round(400 * 16)), 800))]
    y = sg.soundgen(u=u2, **kw2)
    ref = osg(rng=so.RStream(u=u2[0]), **kw2)
    assert y.size == ref.size and rel(y, ref) < TOL


def test_failed_syllable_reports_like_reference():
    with pytest.raises(sg.SoundgenError) as ei:
        sg.generateHarmonics(np.full(3, 100.))
    assert 'Failed to generate the new syllable' in str(ei.value)


def test_linearity_property_full_size():
    # size-independent property at a BASELINE size: scaling a formant-free envelope leaves the
    # normalised output unchanged; the filter of a 1 s 48 kHz sound has the reference's length
    x = np.random.default_rng(0).standard_normal(48000)
    a = sg.filter_sound(x, np.full(1200, 1.0), 2400)
    b = sg.filter_sound(x, np.full(1200, 7.5), 2400)
    assert a.size == 47400 and np.max(np.abs(a - b)) < 1e-5


def test_both_synthesis_kernels_agree():
    """K1 has two kernels (tcgen05 contraction for epochs with many rows, packed-FP32 Clenshaw for the others): the
    same batch through either gives the same integer artefacts and waveforms within 2e-5 of peak, and each agrees
    with the oracle (the dispatch threshold is moved with sgb_synth_min_rows_set)."""
    from soundgen_beta_b200 import _abi
    L = _abi.load()
    calls = workloads.config3(n=24) + workloads.config1(n=8)
    outs = {}
    try:
        for mode, rows in (('tc', 0), ('ffma', 1 << 30), ('default', -1)):
            assert L.sgb_synth_min_rows_set(rows) == 0
            bb = sg.BatchBuilder(u_dtype=np.float32)
            for kw in calls:
                bb.add_soundgen(**dict(kw))
            bt = sg.Batch()
            bt.upload(bb.build())
            bt.run()
            outs[mode] = [np.array(w, copy=True) for w in bt.fetch(np.float64)]
            bt.close()
    finally:
        L.sgb_synth_min_rows_set(-1)
    for a, b, c in zip(outs['tc'], outs['ffma'], outs['default']):
        assert a.shape == b.shape == c.shape        # zero-crossing trims agree: same lengths
        pk = np.max(np.abs(b))
        assert np.max(np.abs(a - b)) <= 2e-5 * pk and np.max(np.abs(c - b)) <= 2e-5 * pk
    for i in (0, 5, 23, 27):
        ref = _oracle_call(calls[i])
        for mode in ('tc', 'ffma'):
            y = outs[mode][i]
            assert y.shape == ref.shape
            assert np.max(np.abs(y - ref)) <= 1e-4 * np.max(np.abs(ref)), (mode, i)


@pytest.mark.parametrize('rows', [7, 16, 17, 127, 128, 129, 383, 384, 385, 767, 769, 1151, 1152, 1153, 1300])
def test_tensor_core_kernel_at_pass_boundaries(rows):
    """k_synth_tc works in blocks of 16 rows, passes of 384 and resident groups of 1152: epochs whose row count sits
    on either side of each boundary (flat pitch at 48 kHz, every harmonic below Nyquist kept) against the oracle,
    with the dispatch forced to the tensor-core kernel."""
    from soundgen_beta_b200 import _abi
    L = _abi.load()
    f0 = 24000.0 / (rows + 0.5)                       # harmonics 1..rows lie below Nyquist
    kw = dict(sylLen=max(300, int(6000.0 / f0)), samplingRate=48000, invalidArgAction='ignore', temperature=0,
              pitchAnchors=[f0, f0 * 1.0001], pitchFloor=1, rolloff=-1, rolloffOct=0, rolloffKHz=0, nonlinBalance=0,
              attackLen=10, addSilence=0)
    ref = _oracle_call(kw)
    try:
        assert L.sgb_synth_min_rows_set(0) == 0
        y = sg.soundgen(**kw)
    finally:
        L.sgb_synth_min_rows_set(-1)
    assert y.shape == ref.shape
    assert np.max(np.abs(y - ref)) <= 1e-4 * np.max(np.abs(ref))
