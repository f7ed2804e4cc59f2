"""Small host-side helpers of the Python mirror of soundgen's R shells (argument shapes, the vowel
dictionary).  The host stage of soundgen() itself -- validation, hyper-parameters, R's random stream,
contours -- lives in the library's C front-end (csrc/frontend.cu).
"""
from __future__ import annotations

import math

import numpy as np

def rint(x):
    return np.rint(np.asarray(x, dtype=np.float64))


def seq_len(frm, to, n):
    """seq(from, to, length.out = n)."""
    n = int(math.ceil(n))
    if n <= 0:
        return np.zeros(0)
    if n == 1:
        return np.array([float(frm)])
    by = (float(to) - float(frm)) / float(n - 1)
    out = float(frm) + np.arange(n, dtype=np.float64) * by
    out[0], out[-1] = float(frm), float(to)
    return out


def as_anchors(a, t_hi=1.0):
    """numeric vector / (time, value) / dict -> (time[], value[]) or None
    (R/soundgen.R:305-315)."""
    if a is None:
        return None
    if isinstance(a, dict):
        return np.asarray(a['time'], dtype=np.float64), np.asarray(a['value'], dtype=np.float64)
    if isinstance(a, (tuple, list)) and len(a) == 2 and np.ndim(a[0]) == 1 and np.ndim(a[1]) == 1 \
            and len(a[0]) == len(a[1]) and not np.isscalar(a[0]):
        return np.asarray(a[0], dtype=np.float64), np.asarray(a[1], dtype=np.float64)
    v = np.atleast_1d(np.asarray(a, dtype=np.float64))
    return seq_len(0, t_hi, v.size), v


# presets$M1$Formants$vowels (R/presets.R:175-213): rows of (freq, amp, width) per named formant
M1_VOWELS = {
    'a': {'f1': (860, 30, 120), 'f2': (1280, 40, 120), 'f3': (2900, 25, 200)},
    'o': {'f1': (630, 35, 100), 'f2': (900, 35, 100), 'f3': (3000, 30, 200), 'f4': (3960, 30, 200)},
    'i': {'f1': (300, 25, 80), 'f2': (2700, 30, 100), 'f3': (3400, 40, 350), 'f4': (4200, 40, 350)},
    'e': {'f1': (530, 30, 50), 'f1.4': (1100, -20, 100), 'f1.6': (1400, 20, 100), 'f2': (2400, 40, 300),
          'f3': (4000, 30, 300)},
    'u': {'f1': (375, 25, 80), 'f2': (550, 35, 120), 'f3': (2100, 25, 300), 'f4': (4200, 45, 250)},
    '0': {'f1': (640, 30, 100), 'f2': (1670, 30, 100), 'f3': (2700, 30, 100), 'f4': (3880, 30, 100)},
}


def convert_string_to_formants(phonemeString, vowels=None):
    """convertStringToFormants (R/utilities_soundgen.R:135-222), speaker 'M1': a string of vowels
    ('aui') -> list of (k, 4) arrays (time, freq, amp, width), k = number of valid phonemes, formants in
    the order of their sorted names.  A formant that a vowel lacks is filled in with amplitude 0 and the
    frequency the first vowel that has it gives it (:170-185).  Returns None where R returns NA."""
    vowels = M1_VOWELS if vowels is None else vowels
    valid = [c for c in phonemeString if c in vowels]
    if not valid:
        return None
    uniq = list(dict.fromkeys(valid))
    names = sorted({f for v in uniq for f in vowels[v]})
    filled = {}
    for v in uniq:
        d = dict(vowels[v])
        for f in names:
            if f not in d:
                closest = [vowels[w][f][0] for w in uniq if f in vowels[w]][0]
                d[f] = (closest, 0, 100)
        filled[v] = d
    times = seq_len(0, 1, len(valid))
    out = []
    for f in names:
        rows = np.array([[times[k], *filled[v][f]] for k, v in enumerate(valid)], dtype=np.float64)
        # :215-218 drops a formant when its number of zero amplitudes equals length(f) == 4 columns
        if np.sum(rows[:, 2] == 0) == 4:
            continue
        out.append(rows)
    return out


def seq_by_count(frm, to, by):
    if frm == to:
        return 1
    return int(math.floor((to - frm) / by + 1e-10)) + 1
