"""Host-side logic behind the C ABI: the front-end's argument munging and batch layout, workloads."""
import ctypes as C

import numpy as np
import pytest

import soundgen_beta_b200 as sg
from soundgen_beta_b200 import _abi, host, workloads
from oracle.rprims import r_seq_len_out


def tables(d):
    """ctypes views of a batch description's struct arrays."""
    def arr(ptr, T, n):
        return C.cast(ptr, C.POINTER(T * max(1, n))).contents if ptr and n else []
    return dict(bouts=arr(d.bouts, _abi.Bout, d.n_bouts), syls=arr(d.syllables, _abi.Syllable, d.n_syllables),
                noises=arr(d.noises, _abi.Noise, d.n_noises), envs=arr(d.envelopes, _abi.Envelope, d.n_envelopes))


def test_host_primitives():
    assert np.array_equal(host.seq_len(1, 7, 13), r_seq_len_out(1, 7, 13))
    f = host.convert_string_to_formants('a')          # soundgen.R:384-386 / utilities_soundgen.R:135-222
    assert [tuple(r[0]) for r in f] == [(0, 860, 30, 120), (0, 1280, 40, 120), (0, 2900, 25, 200)]
    f = host.convert_string_to_formants('aui')
    assert len(f) == 4 and all(r.shape == (3, 4) for r in f) and list(f[0][:, 0]) == [0, .5, 1]
    assert list(f[3][:, 1]) == [4200, 4200, 4200] and list(f[3][:, 2]) == [0, 45, 40]   # 'a' has no f4
    assert host.convert_string_to_formants('xyz') is None


def test_builder_layout():
    bb = sg.BatchBuilder()
    c = bb.add_soundgen(sylLen=200, nSyl=3, pauseLen=100, repeatBout=2, pitchAnchors=[100, 150], temperature=0,
                        noiseAnchors=((-50., 250.), (-30., -10.)), u=[np.zeros(bb.noise_uniform_count(4800, 800))] * 6)
    d = bb.build()
    t = tables(d)
    assert c == 0 and d.n_bouts == 2 and d.n_syllables == 6 and d.n_noises == 6
    b0, b1 = t['bouts'][0], t['bouts'][1]
    assert (b0.syl_begin, b0.syl_end, b1.syl_begin, b1.syl_end) == (0, 3, 3, 6)
    assert b0.lead_silence == 1600 and b0.tail_silence == 0 and b1.lead_silence == 1600 and b1.tail_silence == 1600
    assert t['syls'][0].pause_after == 1616 and t['syls'][2].pause_after == 0   # gap = pauseLen + 1 ms (utilities_soundgen.R:546-549)
    assert t['noises'][0].len == 4800 and t['noises'][0].insertion == 1 - 800   # 50 ms pre-aspiration (soundgen.R:522-528)
    # no draw count depends on the pitch contour here: it is evaluated on the device from the two anchors
    assert d.n_calls == 1 and d.n_pitch == 0 and t['syls'][0].pitch_anchor_n == 2
    assert [y.pitch_off for y in t['syls']] == [700 * i for i in range(6)] and t['syls'][5].pitch_len == 700


def test_range_check_and_unsupported():
    bb = sg.BatchBuilder()
    w = []
    bb.add_soundgen(samplingRate=48000, temperature=0, pitchAnchors=[100, 150], warn=w)
    bb.add_soundgen(samplingRate=48000, temperature=0, pitchAnchors=[100, 150], invalidArgAction='ignore')
    assert any('samplingRate' in m for m in w)
    with pytest.raises(ValueError):
        bb.add_soundgen(sylLen=10, temperature=0, invalidArgAction='abort')
    with pytest.raises(NotImplementedError):
        bb.add_soundgen(temperature=0.1)            # host draws need R's stream: a seed
    bb.add_soundgen(temperature=0.1, seed=3)
    t = tables(bb.build())
    assert t['syls'][0].samplingRate == 16000 and t['syls'][1].samplingRate == 48000
    assert t['syls'][2].temperature == 0.1 and t['syls'][2].z_cap > 0


def test_workloads_are_seeded():
    a, b = workloads.config3(n=3), workloads.config3(n=3)
    assert all(np.array_equal(x['z'][0], y['z'][0]) and x['rolloff'] == y['rolloff'] for x, y in zip(a, b))
    c2 = workloads.config2(n=2)
    assert c2[0]['u'][0].size == workloads.noise_uniform_count(
        int(np.rint((c2[0]['noiseAnchors'][0].max() - c2[0]['noiseAnchors'][0].min()) * 22.05)), 1102)
    assert workloads.config0() == [dict(sylLen=1000, seed=1)]           # the literal reference call
    c4 = workloads.config4(n=34)
    assert c4[33]['seed'] == 33 and sorted(c4[0]) == sorted(c4[33])


def _pool_bytes(d):
    esz = 4 if d.u_is_float else 8
    out = []
    for ptr, n in ((d.pitch, 8 * d.n_pitch), (d.anchors, 16 * d.n_anchors), (d.formants, 32 * d.n_formants),
                   (d.z, 8 * d.n_z), (d.u, esz * d.n_u), (d.pre, 8 * d.n_pre),
                   (d.syllables, C.sizeof(_abi.Syllable) * d.n_syllables), (d.noises, C.sizeof(_abi.Noise) * d.n_noises),
                   (d.bouts, C.sizeof(_abi.Bout) * d.n_bouts), (d.envelopes, C.sizeof(_abi.Envelope) * d.n_envelopes),
                   (d.calls, C.sizeof(_abi.Call) * d.n_calls), (d.formant_index, C.sizeof(_abi.FormantRef) * d.n_formant_refs)):
        out.append(C.string_at(ptr, n) if ptr and n else b'')
    return out


def test_add_many_equals_per_call_add():
    """sgb_frontend_add_many (worker threads) + clear + re-use yield the description the per-call path builds."""
    calls = workloads.config3(n=40) + workloads.config0(n=6, seed=3) + workloads.config2(n=5)
    fe1 = sg.FrontEnd(np.float32)
    for kw in calls:
        fe1.add(**dict(kw))
    d1, n1 = fe1.round_begin()
    aa = sg.ArgArray(calls, np.float32)
    fe2 = sg.FrontEnd(np.float32)
    for _ in range(2):                       # second pass: the cleared handle is reused
        fe2.clear()
        assert fe2.add_many(aa) == 0
        d2, n2 = fe2.round_begin()
        assert n1 == n2 == len(calls)
        assert _pool_bytes(d1) == _pool_bytes(d2)
    bad = sg.ArgArray([dict(sylLen=300, temperature=0), dict(sylLen=-5, temperature=0, invalidArgAction='abort')], np.float32)
    fe2.clear()
    with pytest.raises(ValueError):
        fe2.add_many(bad)
    assert fe2.round_begin()[1] == 0         # nothing of a failed add_many stays registered


def test_parallel_round_begin_equals_serial():
    """round_begin on worker threads (contiguous shards of calls, merged in order) builds the serial description,
    byte for byte -- stochastic presets (R's stream per call), noise, several bouts included."""
    L = _abi.load()
    calls = workloads.config4(n=132) + workloads.config3(n=40) + workloads.config2(n=12) + \
        [dict(sylLen=150, nSyl=2, repeatBout=3, seed=5), dict(sylLen=200, temperature=0, pitchAnchors=[100, 150, 120])]
    aa = sg.ArgArray(calls, np.float32)
    out = []
    try:
        for threads in (1, 5):
            L.sgb_host_set_threads(threads)
            fe = sg.FrontEnd(np.float32)
            fe.add_many(aa)
            d, n = fe.round_begin()
            out.append((n, _pool_bytes(d), fe))
    finally:
        L.sgb_host_set_threads(0)
    assert out[0][0] == out[1][0] == len(calls)
    names = ['pitch', 'anchors', 'formants', 'z', 'u', 'pre', 'syllables', 'noises', 'bouts', 'envelopes', 'calls', 'formant_index']
    diff = [nm for nm, a, b in zip(names, out[0][1], out[1][1]) if a != b]
    assert not diff, diff
