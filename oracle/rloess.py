"""stats::loess for one predictor and a handful of points, restated (TEST INFRASTRUCTURE ONLY).

getSmoothContour (R/smoothContours.R:116-153) calls
`loess(anchors_long ~ time, span = span)` + `predict(l, time)` on 3-10 anchors: family
"gaussian", degree 2, surface "interpolate", cell 0.2, one iteration (all defaults).  The
source of loess (netlib dloess as shipped in R 3.4.0's src/library/stats/src/loessf.f,
loessc.c) is NOT under /root/reference and R cannot run here, so this file restates the
published algorithm (Cleveland, Grosse & Shyu 1992; the dloess routines are named in the
comments).  PARITY UNPINNED -- what is restated, and what could not be checked:

* lowesd: q = min(n, floor(n * span + 1e-5)) nearest points per local fit ("span is too small"
  when q <= 0); fc = floor(n * span * cell) points at most per k-d leaf.
* ehg126: bounding box = data range widened by 0.005 * max(range, 1e-10 * max|x| + 1e-30).
* ehg124/ehg129: a cell with more than fc points is cut THROUGH its m-th smallest point,
  m = (l + u) / 2 (points <= the cut go to the low son); a cell whose m-th point lies on one
  of its own bounds is a leaf; every cut adds a vertex.  With n <= 10 anchors fc is 0 (or 1 for
  long spans), so every anchor (fc = 0) becomes a vertex besides the two box ends and the
  surface interpolates the anchors.  UNCERTAIN: later R versions are believed to cut at the
  midpoint (x[m] + x[m+1]) / 2 instead; the pinned R 3.4.0 is taken to have the older rule,
  which is the only one of the two that terminates cleanly for fc = 0 and that makes the
  reference's default 4-anchor pitch contour pass through its anchors (the midpoint rule makes
  it dip to half the anchor values, below pitchFloor, and the reference's own retry loop
  (smoothContours.R:143-150) would then shrink the span until loess() stops with an error).
* ehg127 at every vertex v: squared distances to v, rho = (q-th smallest) * max(1, span),
  weights sqrt(tricube(sqrt(d2 / rho))), design [1, x - v, (x - v)^2] times the weights, columns
  equilibrated to unit norm, least squares by QR + SVD with singular values below
  100 * eps * sigma_1 dropped (the "pseudoinverse used at ..." warning that soundgen
  suppresses) -- i.e. the minimum-norm solution in the equilibrated coordinates, which is what
  numpy's SVD gives to rounding.  Kept: intercept (fit at v) and slope.
* ehg128: cubic Hermite blend of (value, slope) at the two vertices of the cell containing x.
* predict.loess: NA outside the range of the fitted x.
"""
from __future__ import annotations

import math

import numpy as np


class LoessError(RuntimeError):
    pass


class Loess1D:
    def __init__(self, x, y, span, cell=0.2, degree=2):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        o = np.argsort(x, kind='stable')
        self.x, self.y = x[o], y[o]
        n = x.size
        self.n = n
        self.span = float(span)
        self.q = min(n, int(math.floor(n * span + 1e-5)))
        if self.q <= 0:
            raise LoessError('span is too small')
        self.fc = int(math.floor(n * span * cell))
        self.degree = degree
        lo, hi = float(self.x[0]), float(self.x[-1])
        mu = 0.005 * max(hi - lo, 1e-10 * max(abs(lo), abs(hi)) + 1e-30)
        verts = [lo - mu, hi + mu]
        self._build(0, n - 1, lo - mu, hi + mu, verts)
        self.v = np.array(sorted(set(verts)))
        fits = [self._local_fit(v) for v in self.v]
        self.val = np.array([f[0] for f in fits])
        self.slope = np.array([f[1] for f in fits])

    def _build(self, l, u, vlo, vhi, verts):
        """ehg124/ehg129 for one predictor, 0-based inclusive point range [l, u] of the sorted x."""
        cnt = u - l + 1
        if cnt <= self.fc or not (vhi - vlo > 0):
            return
        m = (l + 1 + u + 1) // 2 - 1          # Fortran m = (l + u) / 2 on 1-based bounds
        # "bug fix from btyner 2006-07-20": move m to the nearest position with x[m] != x[m + 1]
        off = 0
        while True:
            mm = m + off
            if mm >= u or mm < l:
                break
            if self.x[mm] == self.x[mm + 1]:
                off = -off
                if off >= 0:
                    off += 1
                continue
            m = mm
            break
        if self.x[m] == vlo or self.x[m] == vhi:
            return
        t = float(self.x[m])                  # the cut goes through the m-th point
        verts.append(t)
        self._build(l, m, vlo, t, verts)
        self._build(m + 1, u, t, vhi, verts)

    def _local_fit(self, v):
        d2 = (self.x - v) ** 2
        order = np.argsort(d2, kind='stable')[:self.q]
        rho = d2[order[self.q - 1]] * max(1.0, self.span)
        if rho <= 0:
            return float('nan'), float('nan')
        r = np.sqrt(d2[order] / rho)
        w = np.where(r < 1, np.sqrt(np.clip(1 - r ** 3, 0, None) ** 3), 0.0)
        dx = self.x[order] - v
        cols = [w, w * dx]
        if self.degree >= 2:
            cols.append(w * dx * dx)
        B = np.stack(cols, axis=1)
        eta = w * self.y[order]
        nrm = np.sqrt(np.sum(B * B, axis=0))
        nrm[nrm == 0] = 1.0
        B = B / nrm
        U, s, Vt = np.linalg.svd(B, full_matrices=False)
        if s.size == 0 or s[0] == 0:
            return 0.0, 0.0
        tol = s[0] * 100 * np.finfo(np.float64).eps
        coef = np.zeros(B.shape[1])
        for j in range(s.size):
            if s[j] > tol:
                coef += (U[:, j] @ eta) / s[j] * Vt[j]
        coef = coef / nrm
        # a fit through an anchor whose value is exactly 0 is rounding noise of either sign (in R's LINPACK
        # too) and sourceSpectrum.R:443 asks `mouthOpening > 0` of it: noise below 64 eps of the data scale
        # is taken as the exact zero it stands for
        if abs(coef[0]) <= 64 * np.finfo(np.float64).eps * np.max(np.abs(self.y)):
            coef[0] = 0.0
        return float(coef[0]), float(coef[1])

    def ok(self):
        return bool(np.all(np.isfinite(self.val)) and np.all(np.isfinite(self.slope)))

    def predict(self, z):
        z = np.asarray(z, dtype=np.float64)
        if not self.ok():
            raise LoessError('NA/NaN/Inf in foreign function call (arg 5)')
        out = np.full(z.shape, np.nan)
        inside = (z >= self.x[0]) & (z <= self.x[-1])
        zi = z[inside]
        # cell: z <= cut goes to the low son
        c = np.clip(np.searchsorted(self.v, zi, side='left') - 1, 0, self.v.size - 2)
        v0, v1 = self.v[c], self.v[c + 1]
        h = v1 - v0
        u = (zi - v0) / h
        phi0 = (1 - u) ** 2 * (1 + 2 * u)
        phi1 = u ** 2 * (3 - 2 * u)
        psi0 = u * (1 - u) ** 2
        psi1 = -u ** 2 * (1 - u)
        out[inside] = phi0 * self.val[c] + phi1 * self.val[c + 1] + \
            (psi0 * self.slope[c] + psi1 * self.slope[c + 1]) * h
        return out


def smooth_contour_loess(time01, value, length, duration_ms, n_anchors, valueFloor=None):
    """The loess branch of getSmoothContour (R/smoothContours.R:116-153): anchors at `time01`
    (already rescaled to 0..1) on a grid of `length` points.  Returns the contour before the
    final clamping (NaN where predict() gives NA) and the span finally used."""
    length = int(length)
    atp = np.asarray(time01, dtype=np.float64) - np.min(time01)
    atp = atp / np.max(atp) * length
    atp[atp == 0] = 1
    idx = np.trunc(atp).astype(np.int64)            # R truncates a fractional subscript
    long_ = np.full(length, np.nan)
    nz = idx[idx != 0]                              # zero subscripts are dropped ...
    vals = np.asarray(value, dtype=np.float64)
    for k, i in enumerate(nz):                      # ... and the values are used in order
        long_[i - 1] = vals[k % vals.size]
    keep = ~np.isnan(long_)
    px = np.nonzero(keep)[0].astype(np.float64) + 1
    py = long_[keep]
    span = (1 / (1 + math.exp(duration_ms / 500)) + 0.5) / 1.1 ** (n_anchors - 3)
    grid = np.arange(1, length + 1, dtype=np.float64)

    def fit(sp):
        l = Loess1D(px, py, sp)           # "span is too small" propagates, as in the reference
        try:
            return l.predict(grid)
        except LoessError:
            return None
    sc = fit(span)
    while sc is None:
        span = span + 0.1
        sc = fit(span)
    if valueFloor is not None:
        while np.nansum(sc < valueFloor - 1e-6) > 0:
            span = span / 1.1
            sc = fit(span)
            if sc is None:
                raise LoessError('predict() failed while reducing the span')
    return sc, span
