"""Seeded parameter sets shared by the CPU and GPU tests."""
import numpy as np


def voiced_case(seed, temperature=0.0, sr=None, long=False):
    r = np.random.default_rng(seed)
    sr = float(r.choice([16000, 22050, 44100, 48000])) if sr is None else float(sr)
    dur = r.uniform(100, 900) if not long else r.uniform(900, 2000)
    P = int(round(dur * 3.5))
    f0a, f0b = np.exp(r.uniform(np.log(60), np.log(600), 2))
    pitch = np.exp(np.linspace(np.log(f0a), np.log(f0b), P))
    pars = dict(samplingRate=sr, pitchFloor=50, nonlinBalance=float(r.choice([0, 40, 100])),
                jitterDep=r.uniform(0, 3), jitterLen=r.uniform(1, 30), vibratoFreq=r.uniform(3, 8),
                vibratoDep=r.uniform(0, 2), shimmerDep=r.uniform(0, 30), rolloff=r.uniform(-24, -1),
                rolloffOct=r.uniform(-12, 0), rolloffKHz=r.uniform(-12, 0),
                rolloffParab=float(r.choice([0, -20, 15])), rolloffParabHarm=float(r.integers(1, 8)),
                subFreq=r.uniform(25, 150), subDep=float(r.choice([0, r.uniform(20, 150)])),
                shortestEpoch=float(r.choice([50, 100, 300])), temperature=temperature,
                attackLen=float(r.choice([0, 10, 50])))
    z = r.standard_normal(20000)
    anchors = None
    if r.random() < 0.4:
        anchors = (np.array([0., 1.]), r.uniform(40, 120, 2))
    return pitch, z, anchors, pars
