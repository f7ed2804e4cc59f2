"""The scalar control-rate code that runs on the device (csrc/ctrl.cuh), compiled for the host
by tests/hostsim, against the oracle: integer artefacts bit-exact, amplitudes and phase to
rounding.  No GPU needed."""
import numpy as np

import hostsim_util as hu
from cases import voiced_case
from oracle import soundgen_oracle as so


def _lib():
    return hu.build()


def test_control_artefacts_bit_exact():
    L = _lib()
    for seed in range(24):
        for T in (0.0, 0.3):
            pitch, z, anchors, pars = voiced_case(seed, T)
            y, art = so.generateHarmonics(pitch, rng=so.RStream(z=z), amplAnchors=anchors,
                                          want_artefacts=True, **pars)
            rc, C, out = hu.control(L, pitch, z=z, anchors=anchors, **pars)
            assert rc == 0
            G = C.nGC
            assert np.array_equal(out['gc'][:G], art.gc)
            assert np.array_equal(out['gcup'][:G + 1], art.gc_upsampled)
            assert (C.nHarmonics, C.rows_kept, C.n_up, C.z_used) == \
                   (art.nHarmonics, art.rows_kept, art.n_upsampled, art.z_used)
            assert np.array_equal(np.array(C.ep_start[:C.nEpochs]), art.epochs[:, 0])
            assert np.array_equal(np.array(C.ep_end[:C.nEpochs]), art.epochs[:, 1])
            if art.nSubharm is not None and C.vf_active:
                assert np.array_equal(out['nsub'][:G], art.nSubharm.astype(int))
            if art.rw_bin is not None:
                assert np.array_equal(out['rwbin'][:G], art.rw_bin.astype(int))
            if art.jitter_idx is not None:
                assert np.array_equal(out['jidx'][:C.n_jidx], art.jitter_idx)
            assert np.max(np.abs(out['ppg'][:G] - art.pitch_per_gc) / art.pitch_per_gc) < 1e-13


def test_amplitudes_and_phase():
    L = _lib()
    for seed in range(12):
        for T in (0.0, 0.3):
            pitch, z, anchors, pars = voiced_case(seed, T)
            y, art = so.generateHarmonics(pitch, rng=so.RStream(z=z), amplAnchors=anchors,
                                          want_artefacts=True, **pars)
            rc, C, mats = hu.amplitudes(L, pitch, z=z, anchors=anchors, **pars)
            assert rc == 0
            for e, (m, nm) in enumerate(art.mats):
                n = C.ep_nsub[e] if C.vf_active else 0
                dense = np.zeros_like(mats[e])
                jj = np.rint(nm * (n + 1)).astype(int)
                dense[jj - 1, :] = m
                assert np.max(np.abs(dense - mats[e])) <= 1e-12 * np.max(m)
            rc, C, out = hu.control(L, pitch, z=z, anchors=anchors, **pars)
            G, N = C.nGC, C.n_up
            kt = out['kt'][:G]
            u = np.arange(1, N + 1, dtype=np.float64)
            i = np.clip(np.searchsorted(kt, u, side='right') - 1, 0, G - 1)
            M = u - kt[i]
            s1 = M * (M + 1) / 2
            ph = (out['phi'][i] + out['ppg'][i] * (M + 1) + out['sb'][i] * s1 +
                  out['sc'][i] * (M * (M + 1) * (2 * M + 1) / 6) + out['sd'][i] * s1 * s1) / pars['samplingRate']
            assert np.max(np.abs(ph - art.integr)) < 1e-9   # cycles


def test_failure_codes():
    L = _lib()
    rc, C, out = hu.control(L, np.full(3, 100.), samplingRate=16000)   # one glottal cycle: approx() fails in R
    assert rc == -4
