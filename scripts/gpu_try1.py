import sys, time, traceback
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np
import soundgen_beta_b200 as sg
from oracle import soundgen_oracle as so
from oracle.soundgen_call import soundgen as osg
from cases import voiced_case

def rel(a, b):
    if a.shape != b.shape: return 'SHAPE %s vs %s' % (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

def step(name, f):
    t = time.time()
    try:
        print(name, '->', f(), '(%.2fs)' % (time.time() - t), flush=True)
    except Exception as e:
        print(name, 'EXC', repr(e)[:300], flush=True); traceback.print_exc()

step('rolloff', lambda: rel(sg.getRolloff([150, 800, 3000], rolloffOct=0), so.getRolloff([150, 800, 3000], rolloffOct=0)[0]))
fm = [np.array([[0, 860, 30, 120.]]), np.array([[0, 1280, 40, 120.]]), np.array([[0, 2900, 25, 200.]])]
step('env static', lambda: rel(sg.getSpectralEnvelope(400, 1, formants=fm, vocalTract=15.5, mouthAnchors=[.5, .5]),
                              so.getSpectralEnvelope(400, 1, formants=fm, vocalTract=15.5, mouthAnchors=(np.array([0., 1]), np.array([.5, .5])))))
fm2 = [np.array([[0, 860, 30, 120.], [1, 500, 35, 100]]), np.array([[0, 1280, 40, 120.], [1, 2000, 20, 150]])]
step('env moving', lambda: rel(sg.getSpectralEnvelope(551, 40, formants=fm2, vocalTract=15.5, mouthAnchors=[0, .8]),
                              so.getSpectralEnvelope(551, 40, formants=fm2, vocalTract=15.5, mouthAnchors=(np.array([0., 1]), np.array([0, .8])))))
def harm(seed, T):
    pitch, z, anchors, pars = voiced_case(seed, T)
    ref, art = so.generateHarmonics(pitch, rng=so.RStream(z=z), amplAnchors=anchors, want_artefacts=True, **pars)
    y, a = sg.generateHarmonics(pitch, z=z, amplAnchors=anchors, want_artefacts=True, **pars)
    ok = (np.array_equal(a['gc'], art.gc), np.array_equal(a['gc_upsampled'], art.gc_upsampled), np.array_equal(a['epochs'], art.epochs),
          [tuple(int(v or 0) for v in zz) for zz in art.zc] == [tuple(r) for r in a['zc'].tolist()])
    return rel(y, ref), ok, y.size
for seed in range(6):
    for T in (0.0, 0.3):
        step('harm %d T=%g' % (seed, T), lambda: harm(seed, T))
def filt():
    x = np.random.default_rng(3).standard_normal(16000)
    env = so.getSpectralEnvelope(400, 1, formants=fm, vocalTract=15.5)
    return rel(sg.filter_sound(x, env, 800), so.filter_sound(x, env, 800))
step('filter 800', filt)
def filt2(wl, n, moving):
    x = np.random.default_rng(4).standard_normal(n)
    nc = so.frame_starts(n, wl, 75).size
    env = so.getSpectralEnvelope(wl // 2, nc if moving else 1, formants=fm2 if moving else fm, vocalTract=15.5, samplingRate=44100)
    return rel(sg.filter_sound(x, env, wl), so.filter_sound(x, env, wl))
step('filter 2204 moving', lambda: filt2(2204, 22050, True))
step('filter 1102', lambda: filt2(1102, 44100, False))
step('filter 2400', lambda: filt2(2400, 48000, True))
step('filter 160', lambda: filt2(160, 5000, False))
def noise():
    bb = sg.BatchBuilder(); n = bb.noise_uniform_count(8000, 800)
    u = np.random.default_rng(5).random(n)
    an = (np.array([0., 500]), np.array([-20., 10]))
    ref = so.generateNoise(8000, an, rolloffNoise=-6, attackLen=10, windowLength_points=800, rng=so.RStream(u=u))
    return rel(sg.generateNoise(8000, an, rolloffNoise=-6, attackLen=10, windowLength_points=800, u=u), ref)
step('noise', noise)
kw = dict(sylLen=1000, pitchAnchors=[100, 150], temperature=0, addSilence=100)
step('soundgen cfg0', lambda: rel(sg.soundgen(**kw), osg(**kw)))
