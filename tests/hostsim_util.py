"""ctypes access to the TEST-ONLY host build of csrc/ctrl.cuh (tests/hostsim)."""
import ctypes as C
import os
import subprocess
import numpy as np
from soundgen_beta_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
MAXE = 128


class SylCtrl(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ['status', 'nGC', 'nHarmonics', 'rows_kept', 'nEpochs', 'n_up',
                                          'n_jidx', 'z_used', 'parab_harm', 'any_oct', 'vf_active',
                                          'use_ampl', 'out_len', 'tiles', 'tiles_tc', 'pad_tc']] + \
               [('amp_elems', C.c_int64), ('wave_elems', C.c_int64), ('parab_a', C.c_double),
                ('parab_b', C.c_double), ('parab_c', C.c_double), ('raw_max', C.c_double)] + \
               [(n, C.c_int32 * MAXE) for n in ['ep_start', 'ep_end', 'ep_nsub', 'ep_rows', 'ep_zc1', 'ep_zc2']] + \
               [('ep_amp_off', C.c_int64 * MAXE), ('ep_wave_off', C.c_int64 * MAXE)]


def build():
    src = os.path.join(HERE, 'hostsim', 'hostsim.cpp')
    out = os.path.join(HERE, 'hostsim', 'libhostsim.so')
    deps = [src] + [os.path.join(HERE, '..', 'soundgen_beta_b200', 'csrc', f) for f in
                    ('ctrl.cuh', 'rmath.cuh', 'common.cuh')] + [os.path.join(HERE, '..', 'include', 'soundgen_b200.h')]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', out, src])
    L = C.CDLL(out)
    assert L.hs_sizeof_ctrl() == C.sizeof(SylCtrl), (L.hs_sizeof_ctrl(), C.sizeof(SylCtrl))
    assert L.hs_sizeof_syllable() == C.sizeof(_abi.Syllable)
    return L


def make_syllable(pitch_len, z_cap=0, ampl_n=0, **pars):
    s = _abi.Syllable()
    s.kind = 1
    s.pitch_len = pitch_len
    s.z_cap = z_cap
    s.ampl_n = ampl_n
    d = dict(attackLen=50, nonlinBalance=0, jitterDep=0, jitterLen=1, vibratoFreq=100, vibratoDep=0,
             shimmerDep=0, rolloff=-18, rolloffOct=-2, rolloffKHz=-6, rolloffParab=0, rolloffParabHarm=3,
             rolloff_perAmpl=12, temperature=0, pitchDriftDep=.5, pitchDriftFreq=.125,
             randomWalk_trendStrength=.5, shortestEpoch=300, subFreq=100, subDep=0, samplingRate=16000,
             pitchFloor=75, pitchCeiling=3500, pitchSamplingRate=3500, throwaway=-120)
    d.update(pars)
    for k in _abi.SYL_DOUBLES:
        setattr(s, k, float(d[k]))
    return s


def control(L, pitch, z=None, anchors=None, tile=512, **pars):
    pitch = np.ascontiguousarray(pitch, dtype=np.float64)
    z = np.zeros(1) if z is None else np.ascontiguousarray(z, dtype=np.float64)
    an = np.zeros(2) if anchors is None else np.ascontiguousarray(np.stack(anchors, axis=1).ravel(), dtype=np.float64)
    P = pitch.size
    cap = P // 2 + 2
    sr, fl = pars.get('samplingRate', 16000), pars.get('pitchFloor', 75)
    hcap = int(np.ceil((sr / 2 - fl) / fl)) + 1
    s = make_syllable(P, z_cap=z.size, ampl_n=0 if anchors is None else len(anchors[0]), **pars)
    ctrl = SylCtrl()
    ia = lambda n: np.zeros(n, dtype=np.int32)
    da = lambda n: np.zeros(n, dtype=np.float64)
    out = dict(gc=ia(cap), ppg=da(cap), gcup=ia(cap + 1), nsub=ia(cap), rwbin=ia(cap), jidx=ia(cap),
               rowmap=ia(hcap), kt=da(cap), sb=da(cap), sc=da(cap), sd=da(cap), phi=da(cap), drift=da(cap),
               shimmer=da(cap))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.hs_control.argtypes = [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p] * 15
    rc = L.hs_control(C.byref(s), p(pitch), p(an), p(z), cap, hcap, tile, C.byref(ctrl),
                      *[p(out[k]) for k in ['gc', 'ppg', 'gcup', 'nsub', 'rwbin', 'jidx', 'rowmap', 'kt', 'sb',
                                            'sc', 'sd', 'phi', 'drift', 'shimmer']])
    return rc, ctrl, out


def amplitudes(L, pitch, z=None, anchors=None, **pars):
    pitch = np.ascontiguousarray(pitch, dtype=np.float64)
    z = np.zeros(1) if z is None else np.ascontiguousarray(z, dtype=np.float64)
    an = np.zeros(2) if anchors is None else np.ascontiguousarray(np.stack(anchors, axis=1).ravel(), dtype=np.float64)
    P = pitch.size
    cap = P // 2 + 2
    sr, fl = pars.get('samplingRate', 16000), pars.get('pitchFloor', 75)
    hcap = int(np.ceil((sr / 2 - fl) / fl)) + 1
    s = make_syllable(P, z_cap=z.size, ampl_n=0 if anchors is None else len(anchors[0]), **pars)
    ctrl = SylCtrl()
    amp = np.zeros(8_000_000)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.hs_amplitudes.argtypes = [C.c_void_p] * 4 + [C.c_int] * 2 + [C.c_void_p, C.c_void_p, C.c_int64]
    rc = L.hs_amplitudes(C.byref(s), p(pitch), p(an), p(z), cap, hcap, C.byref(ctrl), p(amp), amp.size)
    mats = []
    if rc == 0:
        for e in range(ctrl.nEpochs):
            rows = ctrl.ep_rows[e]
            ge = ctrl.ep_end[e] - ctrl.ep_start[e] + 1
            o = ctrl.ep_amp_off[e]
            mats.append(amp[o:o + rows * ge].reshape(ge, rows).T.copy())
    return rc, ctrl, mats
