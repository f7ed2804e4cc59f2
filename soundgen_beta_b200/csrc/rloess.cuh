// stats::loess for one predictor and at most LOESS_MAXP points: the fit that getSmoothContour
// (R/smoothContours.R:116-153) asks for whenever it is given 3-10 anchors.  Scalar host/device
// code (common.cuh): the host front-end fits pitch contours with it, the kernels fit the contours
// whose length is only known on the device (amplitude envelopes, mouth opening, noise strength).
//
// Defaults of loess(): family gaussian, degree 2, surface "interpolate", cell 0.2.  The R sources
// of loess are not in the reference tree; the algorithm is the published dloess one
// (Cleveland, Grosse & Shyu), routine names in the comments:
//   lowesd  q = min(n, floor(n span + 1e-5)) nearest points per local fit, fc = floor(n span cell)
//   ehg126  bounding box = data range widened by 0.5 %
//   ehg124  cells holding more than fc points are cut through their median point; each cut is a vertex
//   ehg127  at each vertex: tricube weights over the q nearest points (the q-th gets weight 0),
//           weighted quadratic, columns scaled to unit norm, minimum-norm least squares with
//           singular values below 100 eps sigma_1 dropped ("pseudoinverse used at ...")
//   ehg128  cubic Hermite blend of (value, slope) between the two vertices around x
//   predict NA outside the range of the fitted x
#pragma once
#include "rmath.cuh"

#define LOESS_MAXP 10
#define LOESS_MAXV 12

struct LoessFit {
  int nv;                  // vertices, ascending
  int status;              // 0 ok, 1 "span is too small", 2 non-finite vertex values (predict() fails)
  double x0, x1;           // range of the fitted x: predictions outside are NA
  double v[LOESS_MAXV], val[LOESS_MAXV], slope[LOESS_MAXV];
};

// minimum-norm least squares of B (m x 3, column-major b[col][row]) c = eta
SGB_HD void loess_minnorm3(int m, double b[3][LOESS_MAXP], const double *eta, double *coef) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 40; sweep++) {       // one-sided Jacobi SVD
    bool rotated = false;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < m; i++) { al += b[p][i] * b[p][i]; be += b[q][i] * b[q][i]; ga += b[p][i] * b[q][i]; }
        if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
        rotated = true;
        double zeta = (be - al) / (2.0 * ga);
        double t = ((zeta >= 0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < m; i++) {
          double bp = b[p][i], bq = b[q][i];
          b[p][i] = c * bp - s * bq;
          b[q][i] = s * bp + c * bq;
        }
        for (int i = 0; i < 3; i++) {
          double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[3], smax = 0.0;
  for (int j = 0; j < 3; j++) {
    double a = 0;
    for (int i = 0; i < m; i++) a += b[j][i] * b[j][i];
    sig[j] = sqrt(a);
    smax = fmax(smax, sig[j]);
  }
  const double tol = smax * (100.0 * 2.220446049250313e-16);
  coef[0] = coef[1] = coef[2] = 0.0;
  for (int j = 0; j < 3; j++) {
    if (!(sig[j] > tol)) continue;
    double g = 0;
    for (int i = 0; i < m; i++) g += b[j][i] * eta[i];      // sigma_j * (u_j . eta)
    g = g / (sig[j] * sig[j]);
    for (int i = 0; i < 3; i++) coef[i] += g * V[i][j];
  }
}

// ehg127: local quadratic at vertex v over the q nearest of the n points
SGB_HD void loess_local(int n, const double *x, const double *y, int q, double span, double v,
                        double *val, double *slope) {
  double d2[LOESS_MAXP];
  int ord[LOESS_MAXP];
  for (int i = 0; i < n; i++) { d2[i] = (x[i] - v) * (x[i] - v); ord[i] = i; }
  for (int i = 1; i < n; i++) {          // stable insertion sort by distance
    int oi = ord[i], j = i - 1;
    while (j >= 0 && d2[ord[j]] > d2[oi]) { ord[j + 1] = ord[j]; j--; }
    ord[j + 1] = oi;
  }
  double rho = d2[ord[q - 1]] * fmax(1.0, span);
  if (!(rho > 0.0)) { *val = NAN; *slope = NAN; return; }
  double b[3][LOESS_MAXP], eta[LOESS_MAXP], nrm[3];
  for (int k = 0; k < q; k++) {
    int i = ord[k];
    double r = sqrt(d2[i] / rho);
    double w = 0.0;
    if (r < 1.0) { double t = 1.0 - r * r * r; w = sqrt(t * t * t); }
    double dx = x[i] - v;
    b[0][k] = w; b[1][k] = w * dx; b[2][k] = w * dx * dx;
    eta[k] = w * y[i];
  }
  for (int j = 0; j < 3; j++) {
    double s = 0;
    for (int k = 0; k < q; k++) s += b[j][k] * b[j][k];
    s = sqrt(s);
    if (s > 0.0) { for (int k = 0; k < q; k++) b[j][k] = b[j][k] / s; nrm[j] = s; } else nrm[j] = 1.0;
  }
  double coef[3];
  loess_minnorm3(q, b, eta, coef);
  *val = coef[0] / nrm[0];
  *slope = coef[1] / nrm[1];
  // A fit through an anchor whose value is exactly 0 comes out as rounding noise of either sign (in R's
  // LINPACK as well), and getSpectralEnvelope asks `mouthOpening > 0` of it (sourceSpectrum.R:443): noise
  // below 64 eps of the data scale is taken as the exact zero it stands for.
  double ymax = 0.0;
  for (int i = 0; i < n; i++) ymax = fmax(ymax, fabs(y[i]));
  if (fabs(*val) <= 64.0 * 2.220446049250313e-16 * ymax) *val = 0.0;
}

// x ascending and distinct, n <= LOESS_MAXP
SGB_HD void loess_fit(int n, const double *x, const double *y, double span, LoessFit *F) {
  F->nv = 0; F->status = 0;
  F->x0 = x[0]; F->x1 = x[n - 1];
  int q = (int)floor((double)n * span + 1e-5);
  if (q > n) q = n;
  if (q <= 0) { F->status = 1; return; }
  const int fc = (int)floor((double)n * span * 0.2);
  double lo = x[0], hi = x[n - 1];
  double mu = 0.005 * fmax(hi - lo, 1e-10 * fmax(fabs(lo), fabs(hi)) + 1e-30);
  lo -= mu; hi += mu;
  // ehg124 with an explicit stack of cells (point range [l, u], bounds [vlo, vhi])
  bool cut[LOESS_MAXP];
  for (int i = 0; i < n; i++) cut[i] = false;
  int sl[2 * LOESS_MAXP + 2], su[2 * LOESS_MAXP + 2];
  double slo[2 * LOESS_MAXP + 2], shi[2 * LOESS_MAXP + 2];
  int sp = 0;
  sl[0] = 0; su[0] = n - 1; slo[0] = lo; shi[0] = hi; sp = 1;
  while (sp > 0) {
    sp--;
    int l = sl[sp], u = su[sp];
    double vlo = slo[sp], vhi = shi[sp];
    if (u - l + 1 <= fc || !(vhi - vlo > 0.0)) continue;
    int m = (l + u + 2) / 2 - 1;              // Fortran (l + u) / 2 on 1-based bounds
    if (x[m] == vlo || x[m] == vhi) continue;  // x are distinct here: no tie search needed
    cut[m] = true;
    if (sp + 2 > 2 * LOESS_MAXP + 2) break;
    sl[sp] = l; su[sp] = m; slo[sp] = vlo; shi[sp] = x[m]; sp++;
    sl[sp] = m + 1; su[sp] = u; slo[sp] = x[m]; shi[sp] = vhi; sp++;
  }
  int nv = 0;
  F->v[nv++] = lo;
  for (int i = 0; i < n; i++) if (cut[i]) F->v[nv++] = x[i];
  F->v[nv++] = hi;
  F->nv = nv;
  for (int k = 0; k < nv; k++) {
    loess_local(n, x, y, q, span, F->v[k], &F->val[k], &F->slope[k]);
    if (!isfinite(F->val[k]) || !isfinite(F->slope[k])) F->status = 2;
  }
}

// predict(l, z): NAN outside the fitted range
SGB_HD double loess_eval(const LoessFit *F, double z) {
  if (!(z >= F->x0 && z <= F->x1)) return NAN;
  int c = 0;                                  // z <= cut goes to the low cell
  while (c < F->nv - 2 && F->v[c + 1] < z) c++;
  double v0 = F->v[c], v1 = F->v[c + 1], h = v1 - v0;
  double u = (z - v0) / h;
  double phi0 = (1.0 - u) * (1.0 - u) * (1.0 + 2.0 * u);
  double phi1 = u * u * (3.0 - 2.0 * u);
  double psi0 = u * (1.0 - u) * (1.0 - u);
  double psi1 = -(u * u) * (1.0 - u);
  return phi0 * F->val[c] + phi1 * F->val[c + 1] + (psi0 * F->slope[c] + psi1 * F->slope[c + 1]) * h;
}

// sum(predict(l, 1:len) < thr, na.rm = TRUE) > 0 without visiting every grid point: between two
// vertices the surface is one cubic, so its minimum over the integers lies next to a stationary
// point or at the ends of the integer range of the cell.
SGB_HD bool loess_any_below(const LoessFit *F, int len, double thr) {
  for (int c = 0; c + 1 < F->nv; c++) {
    double a = fmax(F->v[c], fmax(F->x0, 1.0)), b = fmin(F->v[c + 1], fmin(F->x1, (double)len));
    double ka = ceil(a), kb = floor(b);
    if (ka > kb) continue;
    double h = F->v[c + 1] - F->v[c];
    double y0 = F->val[c], y1 = F->val[c + 1], s0 = F->slope[c] * h, s1 = F->slope[c + 1] * h;
    // p(u) = y0 + s0 u + (3 (y1 - y0) - 2 s0 - s1) u^2 + (2 (y0 - y1) + s0 + s1) u^3
    double c2 = 3.0 * (y1 - y0) - 2.0 * s0 - s1, c3 = 2.0 * (y0 - y1) + s0 + s1;
    double cand[12];
    int nc = 0;
    cand[nc++] = ka; cand[nc++] = kb;
    double qa = 3.0 * c3, qb = 2.0 * c2, qc = s0;       // p'(u) = qa u^2 + qb u + qc
    double roots[2];
    int nr = 0;
    if (qa != 0.0) {
      double disc = qb * qb - 4.0 * qa * qc;
      if (disc >= 0.0) { double sq = sqrt(disc); roots[nr++] = (-qb - sq) / (2.0 * qa); roots[nr++] = (-qb + sq) / (2.0 * qa); }
    } else if (qb != 0.0) roots[nr++] = -qc / qb;
    for (int r = 0; r < nr; r++) {
      double z = F->v[c] + roots[r] * h;
      for (int d = -1; d <= 2; d++) {
        double k = floor(z) + (double)d;
        if (k >= ka && k <= kb) cand[nc++] = k;
      }
    }
    for (int i = 0; i < nc; i++) {
      double p = loess_eval(F, cand[i]);
      if (p < thr) return true;
    }
  }
  return false;
}

// The loess branch of getSmoothContour (smoothContours.R:120-153): anchors (time already
// rescaled to 0..1, values already clamped / converted) on a grid of len points.
// Returns F->status: 0 ok, 1 loess() stops with "span is too small" (the reference fails too).
SGB_HD void contour_loess_fit(int na, const double *t01, const double *val, int len, double duration_ms,
                              bool has_floor, double valueFloor, LoessFit *F) {
  double px[LOESS_MAXP], py[LOESS_MAXP];
  int np = 0;
  double tmin = t01[0], tmax;
  for (int i = 1; i < na; i++) tmin = fmin(tmin, t01[i]);
  tmax = t01[0] - tmin;
  for (int i = 1; i < na; i++) tmax = fmax(tmax, t01[i] - tmin);
  // anchors_long[anchor_time_points] = anchors$value: a fractional subscript is truncated, zero
  // subscripts are dropped (the values then pair up with the remaining subscripts in order), a
  // later assignment to the same element wins
  int vi = 0;
  for (int i = 0; i < na; i++) {
    double p = (t01[i] - tmin) / tmax * (double)len;
    if (p == 0.0) p = 1.0;
    double idx = trunc(p);
    if (idx == 0.0) continue;
    double v = val[vi % na];
    vi++;
    int j = 0;
    while (j < np && px[j] != idx) j++;
    if (j < np) { py[j] = v; continue; }
    j = np++;
    while (j > 0 && px[j - 1] > idx) { px[j] = px[j - 1]; py[j] = py[j - 1]; j--; }
    px[j] = idx; py[j] = v;
  }
  F->nv = 0; F->status = 1;
  if (np < 1) return;
  double span = (1.0 / (1.0 + exp(duration_ms / 500.0)) + 0.5) / pow(1.1, (double)(na - 3));
  loess_fit(np, px, py, span, F);
  int guard = 0;
  while (F->status == 2 && guard++ < 1000) {     // predict() failed: larger span (:139-143)
    span = span + 0.1;
    loess_fit(np, px, py, span, F);
  }
  if (F->status != 0) return;
  if (has_floor) {
    while (loess_any_below(F, len, valueFloor - 1e-6)) {   // :145-152
      span = span / 1.1;
      loess_fit(np, px, py, span, F);
      if (F->status != 0) { F->status = 1; return; }
    }
  }
}
