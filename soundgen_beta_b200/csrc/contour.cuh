// Device-side getSmoothContour for the anchor counts that need no loess.
#pragma once
#include "rmath.cuh"
#ifndef ENV_MAXK
#define ENV_MAXK 64
#endif

// getSmoothContour(len = n, anchors) for 1, 2 anchors or spline, clamped to [lo, hi]
// (R/smoothContours.R:98-157); anchors interleaved (time, value).
__device__ inline double contour_at(const double *an, int na, int n, int k, double lo, double hi) {
  if (na == 1) return fmin(fmax(an[1], lo), hi);
  double v0 = fmin(fmax(an[1], lo), hi);
  if (na == 2) { double v1 = fmin(fmax(an[3], lo), hi); return r_seq_at(v0, v1, n, k); }
  double tx[ENV_MAXK], vy[ENV_MAXK], b[ENV_MAXK], c[ENV_MAXK], d[ENV_MAXK];
  if (na > ENV_MAXK) na = ENV_MAXK;
  double tmin = an[0], tmax = an[0];
  for (int i = 1; i < na; i++) { tmin = fmin(tmin, an[2 * i]); tmax = fmax(tmax, an[2 * i]); }
  for (int i = 0; i < na; i++) { tx[i] = (an[2 * i] - tmin) / (tmax - tmin); vy[i] = fmin(fmax(an[2 * i + 1], lo), hi); }
  fmm_coef(na, tx, vy, b, c, d);
  double v = r_spline_at(na, tx, vy, b, c, d, n, k);
  return fmin(fmax(v, lo), hi);
}

