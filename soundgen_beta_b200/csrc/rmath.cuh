// Base-R numeric primitives as scalar host/device code (see common.cuh).
// Semantics follow SURVEY.md Appendix A; the R 3.4 C sources are not in the
// reference tree, so the algorithms are restated from their definitions.
#pragma once
#include "common.cuh"

// seq.int(from, to, length.out = n)[k]  (k = 0..n-1): from + k*by, exact ends.
SGB_HD double r_seq_at(double from, double to, int n, int k) {
  if (n <= 1 || k <= 0) return from;
  if (k >= n - 1) return to;
  double by = (to - from) / (double)(n - 1);
  return from + (double)k * by;
}

// the same with the step by = (to - from) / (n - 1) computed once by the caller
SGB_HD double r_seq_by_step(double from, double to, int n) { return (n > 1) ? (to - from) / (double)(n - 1) : 0.0; }
SGB_HD double r_seq_at_by(double from, double to, double by, int n, int k) {
  if (n <= 1 || k <= 0) return from;
  if (k >= n - 1) return to;
  return from + (double)k * by;
}

// Forsythe-Malcolm-Moler cubic spline coefficients (R `spline`, method "fmm").
// x,y: n knots; b,c,d: outputs (also used as workspace).  n >= 2.
SGB_HD void fmm_coef(int n, const double *x, const double *y, double *b, double *c, double *d) {
  if (n < 2) { if (n == 1) { b[0] = c[0] = d[0] = 0.0; } return; }
  if (n < 3) {
    double t = (y[1] - y[0]);
    b[0] = t / (x[1] - x[0]);
    b[1] = b[0];
    c[0] = c[1] = d[0] = d[1] = 0.0;
    return;
  }
  const int nm1 = n - 1;   // index of last knot (0-based)
  // 0-based transcription: X[i] (1-based) == x[i-1]
  d[0] = x[1] - x[0];
  c[1] = (y[1] - y[0]) / d[0];
  for (int i = 1; i < nm1; i++) {
    d[i] = x[i + 1] - x[i];
    b[i] = 2.0 * (d[i - 1] + d[i]);
    c[i + 1] = (y[i + 1] - y[i]) / d[i];
    c[i] = c[i + 1] - c[i];
  }
  b[0] = -d[0];
  b[nm1] = -d[nm1 - 1];
  c[0] = 0.0;
  c[nm1] = 0.0;
  if (n > 3) {
    c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0]);
    c[nm1] = c[nm1 - 1] / (x[nm1] - x[nm1 - 2]) - c[nm1 - 2] / (x[nm1 - 1] - x[nm1 - 3]);
    c[0] = c[0] * d[0] * d[0] / (x[3] - x[0]);
    c[nm1] = -c[nm1] * d[nm1 - 1] * d[nm1 - 1] / (x[nm1] - x[nm1 - 3]);
  }
  for (int i = 1; i <= nm1; i++) {
    double t = d[i - 1] / b[i - 1];
    b[i] = b[i] - t * d[i - 1];
    c[i] = c[i] - t * c[i - 1];
  }
  c[nm1] = c[nm1] / b[nm1];
  for (int i = nm1 - 1; i >= 0; i--) c[i] = (c[i] - d[i] * c[i + 1]) / b[i];
  b[nm1] = (y[nm1] - y[nm1 - 1]) / d[nm1 - 1] + d[nm1 - 1] * (c[nm1 - 1] + 2.0 * c[nm1]);
  for (int i = 0; i < nm1; i++) {
    b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i]);
    d[i] = (c[i + 1] - c[i]) / d[i];
    c[i] = 3.0 * c[i];
  }
  c[nm1] = 3.0 * c[nm1];
  d[nm1] = d[nm1 - 1];
}

// largest i in [0, n-1] with x[i] <= u (0 if u < x[0])
SGB_HD int upper_interval(int n, const double *x, double u) {
  int lo = 0, hi = n;  // invariant: x[lo] <= u (or lo == 0), x[hi] > u (or hi == n)
  while (hi > lo + 1) {
    int k = (lo + hi) >> 1;
    if (u < x[k]) hi = k; else lo = k;
  }
  return lo;
}

SGB_HD double fmm_eval(int n, const double *x, const double *y, const double *b, const double *c,
                       const double *d, double u) {
  int i = upper_interval(n, x, u);
  double dx = u - x[i];
  return y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]));
}

// spline(y, n = nout, x = x)$y[k]: evaluate at seq(x[0], x[n-1], length.out = nout)[k].
// Handles the degenerate single-knot case (constant).
SGB_HD double r_spline_at(int n, const double *x, const double *y, const double *b, const double *c,
                          const double *d, int nout, int k) {
  if (n == 1) return y[0];
  double u = r_seq_at(x[0], x[n - 1], nout, k);
  return fmm_eval(n, x, y, b, c, d, u);
}

SGB_HD double r_spline_at_by(int n, const double *x, const double *y, const double *b, const double *c,
                             const double *d, double by, int nout, int k) {
  if (n == 1) return y[0];
  double u = r_seq_at_by(x[0], x[n - 1], by, nout, k);
  return fmm_eval(n, x, y, b, c, d, u);
}

// approx(y, n = nout, x = x)$y[k] (linear, rule 1).
SGB_HD double r_approx_at(int n, const double *x, const double *y, int nout, int k) {
  double v = r_seq_at(x[0], x[n - 1], nout, k);
  int i = upper_interval(n, x, v);
  if (i > n - 2) i = n - 2;
  int j = i + 1;
  if (v == x[j]) return y[j];
  if (v == x[i]) return y[i];
  return y[i] + (y[j] - y[i]) * ((v - x[i]) / (x[j] - x[i]));
}

// Neumaier-compensated sum: stands in for R's long-double accumulators.
struct CompSum {
  double s, c;
  SGB_HD CompSum() : s(0.0), c(0.0) {}
  SGB_HD void add(double x) {
    double t = s + x;
    if (fabs(s) >= fabs(x)) c += (s - t) + x; else c += (x - t) + s;
    s = t;
  }
  SGB_HD double value() const { return s + c; }
};

// R mean(): long-double mean with one refinement pass.
SGB_HD double r_mean(const double *x, int n) {
  CompSum a;
  for (int i = 0; i < n; i++) a.add(x[i]);
  double m = a.value() / (double)n;
  CompSum t;
  for (int i = 0; i < n; i++) t.add(x[i] - m);
  return m + t.value() / (double)n;
}
