"""Base-R primitives restated in numpy (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference leans on R 3.4 `stats`/`base` C code that is NOT under
/root/reference (packrat/packrat.lock:3 pins R 3.4.0).  Each function below
restates the documented algorithm of one primitive as the hot path uses it;
SURVEY.md Appendix A lists the call sites.  Indices are R's 1-based values held
in float64/int64 arrays unless a docstring says otherwise.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble  # x87 80-bit on x86-64: what R's LDOUBLE accumulators use


def r_round(x):
    """R < 4.0 `round(x)`: nearbyint, i.e. round-half-even on the double.
    Call sites: utilities_soundgen.R:394,408; source.R:274; subharmonics.R:115,129."""
    return np.rint(np.asarray(x, dtype=np.float64))


def r_cumsum(x):
    """R `cumsum`: long-double running sum, each prefix rounded to double
    (source.R:385 `integr`, utilities_soundgen.R:395)."""
    x = np.asarray(x, dtype=np.float64)
    return np.cumsum(x.astype(LD)).astype(np.float64)


def r_sum(x):
    """R `sum` of doubles: long-double accumulation, rounded once at the end."""
    x = np.asarray(x, dtype=np.float64)
    return float(np.sum(x.astype(LD)))


def r_mean(x):
    """R `mean`: long-double mean plus one compensation pass
    (source.R:254 `rw - mean(rw) + 1`, :318 `drift - mean(drift)`)."""
    x = np.asarray(x, dtype=np.float64).astype(LD)
    n = x.size
    s = np.sum(x) / LD(n)
    t = np.sum(x - s)
    s = s + t / LD(n)
    return float(s)


def r_seq_len_out(frm, to, n):
    """R `seq.int(from, to, length.out = n)`: from + k * ((to-from)/(n-1)) with
    both end points exact.  A fractional length.out is rounded up
    (utilities_math.R:648)."""
    n = int(np.ceil(n))
    if n <= 0:
        return np.zeros(0)
    if n == 1:
        return np.array([float(frm)])
    frm = float(frm)
    to = float(to)
    by = (to - frm) / float(n - 1)
    out = frm + np.arange(n, dtype=np.float64) * by
    out[0] = frm
    out[-1] = to
    return out


def r_seq_by(frm, to, by):
    """R `seq(from, to, by)`: from + (0:m)*by, m = floor((to-from)/by + 1e-10)
    (frame starts: soundgen.R:744-746, source.R:88-90, seewave.r:3467)."""
    frm = float(frm)
    to = float(to)
    by = float(by)
    if frm == to:
        return np.array([frm])
    m = int(np.floor((to - frm) / by + 1e-10))
    return frm + np.arange(m + 1, dtype=np.float64) * by


def _regularize(x, y):
    """stats:::regularize.values with ties = mean: sort by x, average ties."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    o = np.argsort(x, kind="stable")
    x = x[o]
    y = y[o]
    if x.size > 1 and np.any(np.diff(x) == 0):
        ux, inv = np.unique(x, return_inverse=True)
        uy = np.zeros(ux.size)
        cnt = np.zeros(ux.size)
        np.add.at(uy, inv, y)
        np.add.at(cnt, inv, 1.0)
        x, y = ux, uy / cnt
    return x, y


def r_approx(y, n, x=None):
    """R `approx(y, n = n, x = x)$y` (linear, rule = 1).

    Evaluates at seq(x[1], x[nx], length.out = n); interval by the C routine's
    bisection (largest i with x[i] <= v, with the exact-knot shortcuts) and the
    association y[i] + (y[j]-y[i]) * ((v-x[i])/(x[j]-x[i])).
    Call sites: source.R:403-405,460-462; utilities_math.R:316;
    sourceSpectrum.R:331-332."""
    y = np.asarray(y, dtype=np.float64)
    if x is None:
        x = np.arange(1, y.size + 1, dtype=np.float64)
    x, y = _regularize(x, y)
    if x.size < 2:
        raise ValueError("need at least two non-NA values to interpolate")
    xout = r_seq_len_out(x[0], x[-1], n)
    nx = x.size
    i = np.searchsorted(x, xout, side="right") - 1  # largest i with x[i] <= v
    i = np.clip(i, 0, nx - 2)
    j = i + 1
    xi, xj, yi, yj = x[i], x[j], y[i], y[j]
    out = yi + (yj - yi) * ((xout - xi) / (xj - xi))
    out = np.where(xout == xj, yj, out)
    out = np.where(xout == xi, yi, out)
    return out


def fmm_coef(x, y):
    """Forsythe-Malcolm-Moler cubic spline coefficients (R `spline_coef`,
    method "fmm").  Returns (b, c, d) such that on [x_i, x_{i+1}]
    s(u) = y_i + dx*(b_i + dx*(c_i + dx*d_i)), dx = u - x_i."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    b = np.zeros(n)
    c = np.zeros(n)
    d = np.zeros(n)
    if n < 2:
        raise ValueError("spline needs at least 2 knots")
    if n < 3:
        t = y[1] - y[0]
        b[0] = t / (x[1] - x[0])
        b[1] = b[0]
        return b, c, d
    # 1-based views to mirror the textbook routine
    X = np.concatenate(([0.0], x))
    Y = np.concatenate(([0.0], y))
    B = np.zeros(n + 1)
    C = np.zeros(n + 2)
    D = np.zeros(n + 1)
    nm1 = n - 1
    D[1] = X[2] - X[1]
    C[2] = (Y[2] - Y[1]) / D[1]
    for i in range(2, n):
        D[i] = X[i + 1] - X[i]
        B[i] = 2.0 * (D[i - 1] + D[i])
        C[i + 1] = (Y[i + 1] - Y[i]) / D[i]
        C[i] = C[i + 1] - C[i]
    B[1] = -D[1]
    B[n] = -D[nm1]
    C[1] = 0.0
    C[n] = 0.0
    if n > 3:
        C[1] = C[3] / (X[4] - X[2]) - C[2] / (X[3] - X[1])
        C[n] = C[nm1] / (X[n] - X[n - 2]) - C[n - 2] / (X[nm1] - X[n - 3])
        C[1] = C[1] * D[1] * D[1] / (X[4] - X[1])
        C[n] = -C[n] * D[nm1] * D[nm1] / (X[n] - X[n - 3])
    for i in range(2, n + 1):
        t = D[i - 1] / B[i - 1]
        B[i] = B[i] - t * D[i - 1]
        C[i] = C[i] - t * C[i - 1]
    C[n] = C[n] / B[n]
    for i in range(nm1, 0, -1):
        C[i] = (C[i] - D[i] * C[i + 1]) / B[i]
    B[n] = (Y[n] - Y[nm1]) / D[nm1] + D[nm1] * (C[nm1] + 2.0 * C[n])
    for i in range(1, nm1 + 1):
        B[i] = (Y[i + 1] - Y[i]) / D[i] - D[i] * (C[i + 1] + 2.0 * C[i])
        D[i] = (C[i + 1] - C[i]) / D[i]
        C[i] = 3.0 * C[i]
    C[n] = 3.0 * C[n]
    D[n] = D[nm1]
    return B[1:n + 1].copy(), C[1:n + 1].copy(), D[1:n + 1].copy()


def fmm_eval(x, y, b, c, d, u):
    """R `spline_eval` for method fmm: interval = largest i with x[i] <= u
    (clamped to [0, n-1]); the C code keeps the previous interval while
    x[i] <= u <= x[i+1], which differs only when u hits a knot exactly, where
    both polynomials agree to rounding."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    i = np.searchsorted(x, u, side="right") - 1
    i = np.clip(i, 0, x.size - 1)
    dx = u - x[i]
    return y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]))


def r_spline(y, n, x=None):
    """R `spline(y, n = n, x = x)$y`, method "fmm", evaluated at
    seq(min(x), max(x), length.out = n).  With x omitted, knots are 1..len(y).
    Call sites: source.R:285; utilities_soundgen.R:410-412;
    utilities_math.R:318; sourceSpectrum.R:331; smoothContours.R:117."""
    y = np.asarray(y, dtype=np.float64)
    if x is None:
        x = np.arange(1, y.size + 1, dtype=np.float64)
    x, y = _regularize(x, y)
    b, c, d = fmm_coef(x, y)
    xout = r_seq_len_out(x[0], x[-1], n)
    return fmm_eval(x, y, b, c, d, xout)


def hamming_w(n):
    """seewave hamming.w (seewave.r:7431-7437): symmetric, n-1 in the denominator."""
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2 * np.pi * k / (n - 1))


def hanning_w(n):
    """seewave hanning.w (seewave.r:7444-7450)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2 * np.pi * k / (n - 1))
