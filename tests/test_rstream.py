"""R's random stream and loess, restated twice -- oracle/rrng.py + oracle/rloess.py (numpy) and the
library's host front-end (csrc/rrng.h, rloess.cuh, frontend.cu, through the C ABI) -- pinned by the widely
published R answers and checked against each other.  No GPU needed: the front-end never computes samples."""
import ctypes as C

import numpy as np
import pytest

import soundgen_beta_b200 as sg
from soundgen_beta_b200 import _abi
from oracle import soundgen_oracle as so
from oracle.rrng import RRng, revsort
from oracle.soundgen_call import soundgen as osg

# published R answers (R < 3.6 sample()): set.seed(s); runif / rnorm / rexp / sample(1:10)
RUNIF = {1: [0.2655087, 0.3721239, 0.5728534, 0.9082078, 0.2016819], 42: [0.9148060, 0.9370754, 0.2861395],
         123: [0.2875775, 0.7883051, 0.4089769]}
RNORM = {1: [-0.6264538, 0.1836433, -0.8356286, 1.5952808, 0.3295078, -0.8204684],
         42: [1.37095845, -0.56469817, 0.36312841, 0.63286260, 0.40426832], 123: [-0.56047565, -0.23017749, 1.55870831]}
REXP = {1: [0.7551818, 1.1816428, 0.1457067]}
SAMPLE10 = {1: [3, 4, 5, 7, 2, 8, 9, 6, 10, 1], 123: [3, 8, 4, 7, 6, 1, 10, 9, 2, 5], 42: [10, 9, 3, 6, 4, 8, 5, 1, 2, 7]}


def lib_draw(seed, kind, n, p1=0.0, p2=0.0, skip=0):
    out = np.zeros(n)
    assert _abi.load().sgb_rng_draw(seed, kind, p1, p2, skip, out.ctypes.data, n) == 0
    return out


def test_published_r_answers_oracle_and_library():
    for s, v in RUNIF.items():
        assert np.allclose(RRng(s).runif(len(v)), v, atol=5e-8) and np.allclose(lib_draw(s, 0, len(v)), v, atol=5e-8)
    for s, v in RNORM.items():
        assert np.allclose(RRng(s).rnorm(len(v)), v, atol=5e-8) and np.allclose(lib_draw(s, 1, len(v)), v, atol=5e-8)
    for s, v in REXP.items():
        assert np.allclose(RRng(s).rexp(len(v)), v, atol=5e-8) and np.allclose(lib_draw(s, 2, len(v)), v, atol=5e-8)
    for s, v in SAMPLE10.items():        # sample(1:10): y[i] = x[j], x[j] = x[--n] with j = floor(n * unif_rand())
        r, x, m, got = RRng(s), list(range(1, 11)), 10, []
        for _ in range(10):
            j = r.unif_index(float(m)); got.append(x[j]); m -= 1; x[j] = x[m]
        assert got == v
        u = lib_draw(s, 0, 10)
        x, got = list(range(1, 11)), []
        for i in range(10):
            j = int(np.floor((10 - i) * u[i])); got.append(x[j]); x[j] = x[9 - i]
        assert got == v
    assert [RRng(1).rbinom1(1, .5) for _ in range(1)] == [0] and list(lib_draw(1, 4, 5, 1, .5)) == [0, 0, 1, 1, 0]


def test_library_stream_equals_oracle_stream():
    for seed in (0, 7, 2026):
        assert np.array_equal(lib_draw(seed, 0, 2000), RRng(seed).runif(2000))       # crosses a twist of the state
        assert np.array_equal(lib_draw(seed, 1, 500), RRng(seed).rnorm(500))
        for shape, rate in ((0.3, 1.0), (1.0, 2.0), (2.5, 1.3), (9.0, 0.5), (40.0, 7.0), (1600.0, 40.0)):
            assert np.array_equal(lib_draw(seed, 3, 200, shape, rate), RRng(seed).rgamma(200, shape, rate))
        r = RRng(seed)
        assert np.array_equal(lib_draw(seed, 5, 300, 7), [r.sample_int1(7) for _ in range(300)])
    g = RRng(3).rgamma(20000, 4.0, 2.0)
    assert abs(g.mean() - 2.0) < 0.03 and abs(g.var() - 1.0) < 0.05       # moments: shape / rate, shape / rate^2


def test_revsort_ties_as_sample_prob_sees_them():
    # wiggleAnchors: sample(c('nothing', 'remove', 'add'), prob = c(1 - T, T / 2, T / 2)): the tie is ordered by
    # R's heapsort, not stably -- 'add' comes before 'remove' in the cumulative table
    p, perm = [0.9, 0.05, 0.05], [1, 2, 3]
    revsort(p, perm)
    assert perm == [1, 3, 2]


def contour(t, v, n, sr=16000, lo=None, hi=None, pitch=False, method=0):
    t = np.ascontiguousarray(t, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(n)
    rc = _abi.load().sgb_smooth_contour(t.ctypes.data, v.ctypes.data, v.size, n, float(sr), int(lo is not None), float(lo or 0),
                                        int(hi is not None), float(hi or 0), int(pitch), method, out.ctypes.data)
    return rc, out


def test_loess_contours_library_vs_oracle():
    r = np.random.default_rng(0)
    worst = 0.0
    for it in range(200):
        n = int(r.integers(3, 11))
        length = int(r.choice([7, 50, 350, 1050, 3500, 17500]))
        t = np.sort(r.uniform(0, 1, n)); t[0], t[-1] = 0, 1
        pitch = bool(it % 2)
        v = r.uniform(60, 800, n) if pitch else r.uniform(-100, 30, n)
        lo, hi = (50, 3500) if pitch else (-120, 40)
        try:
            ref = so.getSmoothContour((t, v), length=length, samplingRate=3500, valueFloor=lo, valueCeiling=hi, thisIsPitch=pitch)
        except Exception:    # loess() stops ("span is too small"): the library must fail as well
            assert contour(t, v, length, 3500, lo, hi, pitch)[0] == _abi.SGB_ERR_SYNTH
            continue
        rc, out = contour(t, v, length, 3500, lo, hi, pitch)
        assert rc == 0
        worst = max(worst, float(np.max(np.abs(out - ref) / np.maximum(1, np.abs(ref)))))
    assert worst < 1e-10, worst


def test_loess_passes_through_its_anchors():
    # k-d vertices sit on the anchors and at most three points carry weight in a local quadratic: exact interpolation
    t, v = np.array([0, .1, .9, 1]), np.array([100., 150, 135, 100])
    rc, out = contour(t, v, 1050, 3500, 50, 3500, True)
    assert rc == 0 and np.allclose(out[[0, 104, 944, 1049]], v, rtol=1e-12)


def front_end_state(kw, seed):
    fe = sg.FrontEnd()
    fe.add(seed=seed, **kw)
    d, n = fe.round_begin()
    pitch = np.ctypeslib.as_array(C.cast(d.pitch, C.POINTER(C.c_double)), shape=(max(1, d.n_pitch),)).copy()
    return fe.rng_state(0), pitch, fe.status()[0], d


@pytest.mark.parametrize('kw,seed', [
    (dict(sylLen=300), 1), (dict(sylLen=300), 2), (dict(sylLen=1000), 3),
    (dict(sylLen=200, nSyl=3, pauseLen=80, temperature=.1, nonlinBalance=60, shimmerDep=10,
          noiseAnchors=((0, 100, 250), (-40, -10, -60)), amplAnchors=((0, .5, 1), (120, 60, 120))), 7),
    (dict(sylLen=150, nSyl=2.5, repeatBout=1.5, temperature=.2, formants=None, vocalTract=12), 11),
    (dict(sylLen=120, nSyl=2, temperature=.3, creakyBreathy=-.5, formantsNoise=[np.array([[0, 900, 30, 100.]])],
          noiseAnchors=((0, 120), (-20, -30))), 5)])
def test_front_end_consumes_the_stream_like_the_oracle(kw, seed):
    """After the host stage of one call the library's stream must stand exactly where the oracle's does after
    the whole call: same draws, same order, same counts (incl. the normals the device will consume)."""
    state, pitch, status, d = front_end_state(kw, seed)
    rng = RRng(seed)
    y, arts, _ = osg(rng=rng, want_artefacts=True, **kw)
    assert status == 0
    assert np.array_equal(np.array(rng.state(), dtype=np.uint32).astype(np.int32), state)
    assert d.n_syllables >= len(arts)
