// getSmoothContour (R/smoothContours.R:53-227) as a two-step scalar host/device routine:
// contour_prepare() turns the anchors into a small table once (one thread, or the host front-end),
// contour_eval() evaluates element k of the contour from the table (any thread).
//   1 anchor: flat; 2 anchors: seq(); 3-10 anchors: loess (rloess.cuh) unless method 'spline';
//   more than 10 anchors: FMM spline.
#pragma once
#include "rloess.cuh"
#ifndef ENV_MAXK
#define ENV_MAXK 64
#endif

#define SGB_CONTOUR_LOESS 0    // the reference's default method
#define SGB_CONTOUR_SPLINE 1

struct ContourTab {
  int kind;        // 0 flat, 1 seq between two anchors, 2 FMM spline, 3 loess
  int n;           // spline knots
  int pitch;       // thisIsPitch: smooth on the semitone scale, return Hz
  int status;      // SGB_OK or SGB_ERR_*
  int has_lo, has_hi;
  double lo, hi;   // valueFloor / valueCeiling on the smoothing scale
  double v0, v1;
  double x[ENV_MAXK], y[ENV_MAXK], b[ENV_MAXK], c[ENV_MAXK], d[ENV_MAXK];
  LoessFit L;
};

SGB_HD double hz_to_semitones(double h) { return log2(h / 16.3516) * 12.0; }       // utilities_math.R:16-18
SGB_HD double semitones_to_hz(double s) { return 16.3516 * exp2(s / 12.0); }       // utilities_math.R:25-27

// an: interleaved (time, value) anchors.  len: points of the contour; samplingRate only enters the
// loess span heuristic (duration_ms = len / samplingRate * 1000, smoothContours.R:99,127).
SGB_HD void contour_prepare(ContourTab *T, const double *an, int na, int len, double samplingRate,
                            bool has_lo, double lo, bool has_hi, double hi, bool pitch, int method) {
  T->status = SGB_OK; T->pitch = pitch ? 1 : 0; T->n = 0; T->kind = 0;
  T->has_lo = has_lo ? 1 : 0; T->has_hi = has_hi ? 1 : 0;
  T->v0 = T->v1 = 0.0;
  if (na < 1 || len < 1) { T->status = SGB_ERR_INVALID; return; }
  if (na > ENV_MAXK) { T->status = SGB_ERR_UNSUPPORTED; return; }
  if (na > 10 && method == SGB_CONTOUR_LOESS) method = SGB_CONTOUR_SPLINE;          // :73-76
  double tmin = an[0];
  for (int i = 1; i < na; i++) tmin = fmin(tmin, an[2 * i]);
  double tmax = an[0] - tmin;
  for (int i = 1; i < na; i++) tmax = fmax(tmax, an[2 * i] - tmin);
  for (int i = 0; i < na; i++) {
    double v = an[2 * i + 1];
    if (has_lo && v < lo) v = lo;                                                    // :78-83
    if (has_hi && v > hi) v = hi;
    if (pitch) v = hz_to_semitones(v);
    T->y[i] = v;
    T->x[i] = (an[2 * i] - tmin) / tmax;                                             // :96-98
  }
  if (pitch) { if (has_lo) lo = hz_to_semitones(lo); if (has_hi) hi = hz_to_semitones(hi); }
  T->lo = lo; T->hi = hi;
  if (na == 1) { T->kind = 0; T->v0 = T->y[0]; return; }
  if (na == 2) { T->kind = 1; T->v0 = T->y[0]; T->v1 = T->y[1]; return; }
  if (method == SGB_CONTOUR_SPLINE) {
    // spline() sorts its knots by x
    for (int i = 1; i < na; i++) {
      double xi = T->x[i], yi = T->y[i];
      int j = i - 1;
      while (j >= 0 && T->x[j] > xi) { T->x[j + 1] = T->x[j]; T->y[j + 1] = T->y[j]; j--; }
      T->x[j + 1] = xi; T->y[j + 1] = yi;
    }
    T->kind = 2; T->n = na;
    fmm_coef(na, T->x, T->y, T->b, T->c, T->d);
    return;
  }
  T->kind = 3;
  const double duration_ms = (double)len / samplingRate * 1000.0;
  contour_loess_fit(na, T->x, T->y, len, duration_ms, has_lo, lo, &T->L);
  if (T->L.status != 0) T->status = SGB_ERR_SYNTH;        // loess() stops: the reference call fails
}

// element k (0-based) of the contour of length len
SGB_HD double contour_eval(const ContourTab *T, int len, int k) {
  double v;
  if (T->kind == 0) v = T->v0;
  else if (T->kind == 1) v = r_seq_at(T->v0, T->v1, len, k);
  else {
    if (T->kind == 2) v = r_spline_at(T->n, T->x, T->y, T->b, T->c, T->d, len, k);
    else v = loess_eval(&T->L, (double)(k + 1));
    if (T->has_lo && v < T->lo) v = T->lo;                                           // :155-156
    if (T->has_hi && v > T->hi) v = T->hi;
    if (v != v) v = 0.0;                                                             // NA -> 0 (:224)
  }
  return T->pitch ? semitones_to_hz(v) : v;
}
