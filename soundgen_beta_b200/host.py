"""Host-side helpers of the Python mirror of soundgen's R shells.

In the R deployment these steps stay in R (argument munging, anchors -> contours,
`permittedValues`); here they are restated so that the same C ABI can be driven from
Python.  Nothing in this module touches samples: it only prepares the small control
inputs (pitch contour per syllable, anchors, formant tables) that the library consumes.
"""
from __future__ import annotations

import math

import numpy as np

# R/presets.R:22-79  (default, low, high)
PERMITTED = {
    'repeatBout': (1, 1, 20), 'nSyl': (1, 1, 10), 'sylLen': (300, 20, 5000),
    'pauseLen': (200, 20, 1000), 'temperature': (.025, 0, 1), 'maleFemale': (0, -1, 1),
    'creakyBreathy': (0, -1, 1), 'nonlinBalance': (0, 0, 100), 'nonlinDep': (50, 0, 100),
    'jitterDep': (3, 0, 24), 'jitterLen': (1, 1, 100), 'vibratoFreq': (5, 3, 10),
    'vibratoDep': (0, 0, 3), 'shimmerDep': (0, 0, 100), 'attackLen': (50, 0, 200),
    'rolloff': (-12, -60, 0), 'rolloffOct': (-12, -30, 10), 'rolloffParab': (0, -50, 50),
    'rolloffParabHarm': (3, 1, 20), 'rolloffKHz': (-6, -20, 0), 'rolloffLip': (6, 0, 20),
    'formantDep': (1, 0, 5), 'formantDepStoch': (30, 0, 60), 'vocalTract': (15.5, 2, 100),
    'subFreq': (100, 10, 1000), 'subDep': (100, 0, 500), 'shortestEpoch': (300, 50, 500),
    'amDep': (0, 0, 100), 'amFreq': (30, 10, 100), 'amShape': (0, -1, 1),
    'samplingRate': (16000, 8000, 44100), 'windowLength': (40, 5, 100),
    'rolloffNoise': (-14, -20, 20),
}
NOISE_AMPL = (-120.0, 40.0)
SYLLEN_LOW = 20.0


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


def _fmm(x, y):
    n = x.size
    b = np.zeros(n); c = np.zeros(n); d = np.zeros(n)
    if n < 3:
        b[:] = (y[1] - y[0]) / (x[1] - x[0])
        return b, c, d
    d[0] = x[1] - x[0]
    c[1] = (y[1] - y[0]) / d[0]
    for i in range(1, n - 1):
        d[i] = x[i + 1] - x[i]
        b[i] = 2.0 * (d[i - 1] + d[i])
        c[i + 1] = (y[i + 1] - y[i]) / d[i]
        c[i] = c[i + 1] - c[i]
    b[0] = -d[0]; b[n - 1] = -d[n - 2]
    c[0] = 0.0; c[n - 1] = 0.0
    if n > 3:
        c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
        c[n - 1] = c[n - 2] / (x[n - 1] - x[n - 3]) - c[n - 3] / (x[n - 2] - x[n - 4])
        c[0] = c[0] * d[0] * d[0] / (x[3] - x[0])
        c[n - 1] = -c[n - 1] * d[n - 2] * d[n - 2] / (x[n - 1] - x[n - 4])
    for i in range(1, n):
        t = d[i - 1] / b[i - 1]
        b[i] = b[i] - t * d[i - 1]
        c[i] = c[i] - t * c[i - 1]
    c[n - 1] = c[n - 1] / b[n - 1]
    for i in range(n - 2, -1, -1):
        c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
    b[n - 1] = (y[n - 1] - y[n - 2]) / d[n - 2] + d[n - 2] * (c[n - 2] + 2.0 * c[n - 1])
    for i in range(n - 1):
        b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i])
        d[i] = (c[i + 1] - c[i]) / d[i]
        c[i] = 3.0 * c[i]
    c[n - 1] = 3.0 * c[n - 1]
    d[n - 1] = d[n - 2]
    return b, c, d


def spline(y, n, x=None):
    """stats::spline(y, n = n, x = x)$y, method 'fmm'."""
    y = np.asarray(y, dtype=np.float64)
    x = np.arange(1, y.size + 1, dtype=np.float64) if x is None else np.asarray(x, dtype=np.float64)
    o = np.argsort(x, kind='stable')
    x, y = x[o], y[o]
    b, c, d = _fmm(x, y)
    u = seq_len(x[0], x[-1], n)
    i = np.clip(np.searchsorted(x, u, side='right') - 1, 0, x.size - 1)
    dx = u - x[i]
    return y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]))


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


def smooth_contour(anchors, length, thisIsPitch=False, method='loess', valueFloor=None,
                   valueCeiling=None):
    """getSmoothContour (R/smoothContours.R:53-227) with `len` given.  The loess branch
    (3-10 anchors, R's default) is R-side code that is not available here."""
    time = np.array(anchors[0], dtype=np.float64)
    value = np.array(anchors[1], dtype=np.float64)
    n = time.size
    if n > 10 and method == 'loess':
        method = 'spline'
    if valueFloor is not None:
        value[value < valueFloor] = valueFloor
    if valueCeiling is not None:
        value[value > valueCeiling] = valueCeiling
    if thisIsPitch:
        value = np.log2(value / 16.3516) * 12
        valueFloor = None if valueFloor is None else math.log2(valueFloor / 16.3516) * 12
        valueCeiling = None if valueCeiling is None else math.log2(valueCeiling / 16.3516) * 12
    time = time - np.min(time)
    if n > 1:
        time = time / np.max(time)
    length = int(length)
    if length <= 0:
        return None
    if n == 1:
        sc = np.full(length, value[0])
    elif n == 2:
        sc = seq_len(value[0], value[1], length)
    else:
        if method != 'spline':
            raise NotImplementedError(
                'a contour with 3-10 anchors uses stats::loess in the reference (host-side R); '
                "pass contour_method='spline', 1-2 or >10 anchors, or a pre-evaluated contour")
        sc = spline(value, length, x=time)
        if valueFloor is not None:
            sc[sc < valueFloor] = valueFloor
        if valueCeiling is not None:
            sc[sc > valueCeiling] = valueCeiling
    if thisIsPitch:
        sc = 16.3516 * 2 ** (sc / 12)
    return sc


def seq_by_count(frm, to, by):
    if frm == to:
        return 1
    return int(math.floor((to - frm) / by + 1e-10)) + 1
