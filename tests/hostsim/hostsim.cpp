// TEST-ONLY host build of the scalar control-rate code (csrc/ctrl.cuh).  Lets the
// CPU test-suite check the device-side logic against the oracle without a GPU.
// Never linked into libsoundgen_b200.so.
#include <vector>
#include <cstring>
#include "../../soundgen_beta_b200/csrc/ctrl.cuh"

extern "C" {

struct HsOut {
  SylCtrl C;
};

// Runs K0 for one syllable.  Returns SylCtrl fields through `ctrl_out` (raw struct)
// and per-gc arrays through the pointers (each with capacity cap / cap+1).
int hs_control(const sgb_syllable *sp, const double *pitch, const double *anchors, const double *z,
               int cap, int hcap, int tile, SylCtrl *ctrl_out, int32_t *gc, double *ppg, int32_t *gcup,
               int32_t *nsub, int32_t *rwbin, int32_t *jidx, int32_t *rowmap, double *kt,
               double *sb, double *sc, double *sd, double *phi, double *drift, double *shimmer) {
  int P = sp->pitch_len;
  std::vector<double> pw(P), buf((size_t)cap * 16);
  SylArrays A;
  memset(&A, 0, sizeof(A));
  A.pitch = pw.data();
  A.gc = gc; A.ppg = ppg; A.gcup = gcup; A.nsub = nsub; A.rwbin = rwbin; A.jidx = jidx;
  A.rowmap = rowmap; A.kt = kt; A.sb = sb; A.sc = sc; A.sd = sd; A.phi = phi; A.drift = drift;
  A.shimmer = shimmer;
  double *b = buf.data();
  A.rw = b; b += cap; A.ro = b; b += cap; A.roct = b; b += cap; A.rk = b; b += cap;
  A.subdep = b; b += cap; A.colmax = b; b += cap; A.t1 = b; b += cap; A.t2 = b; b += cap;
  A.t3 = b; b += cap; A.t4 = b; b += cap;
  A.cap = cap; A.hcap = hcap;
  sgb_syllable s = *sp;
  s.z_off = 0; s.pitch_off = 0; s.ampl_off = 0;
  for (int i = 1; i <= P; i++) pw[i - 1] = ctrl_vibrato(s, i, pitch[i - 1]);
  SylCtrl &C = *ctrl_out;
  ctrl_sequential(s, anchors, z, A, C);
  if (C.status != SGB_OK) return C.status;
  for (int g = 0; g < C.nGC; g++) A.colmax[g] = ctrl_colmax(s, A, C, g);
  int kept = 0;
  for (int h = 1; h <= C.nHarmonics; h++) if (ctrl_rowkept(s, A, C, h)) A.rowmap[kept++] = h;
  C.rows_kept = kept;
  ctrl_sizes(A, C, tile);
  return SGB_OK;
}

// Same as hs_control, then fills the dense amplitude matrices (column-major per epoch,
// rows = ep_rows[e]) into amp (capacity amp_cap doubles).
int hs_amplitudes(const sgb_syllable *sp, const double *pitch, const double *anchors, const double *z,
                  int cap, int hcap, SylCtrl *ctrl_out, double *amp, int64_t amp_cap) {
  int P = sp->pitch_len;
  std::vector<double> pw(P), buf((size_t)cap * 24);
  std::vector<int32_t> ib((size_t)(cap + 1) * 6 + hcap);
  SylArrays A;
  memset(&A, 0, sizeof(A));
  A.pitch = pw.data();
  double *b = buf.data();
  double **dp[] = {&A.ppg, &A.rw, &A.ro, &A.roct, &A.rk, &A.shimmer, &A.drift, &A.subdep, &A.colmax,
                   &A.kt, &A.sb, &A.sc, &A.sd, &A.phi, &A.t1, &A.t2, &A.t3, &A.t4};
  for (auto p : dp) { *p = b; b += cap; }
  int32_t *ip = ib.data();
  A.gc = ip; ip += cap + 1; A.nsub = ip; ip += cap + 1; A.rwbin = ip; ip += cap + 1;
  A.jidx = ip; ip += cap + 1; A.gcup = ip; ip += cap + 1; A.rowmap = ip;
  A.cap = cap; A.hcap = hcap;
  sgb_syllable s = *sp;
  s.z_off = 0; s.pitch_off = 0; s.ampl_off = 0;
  for (int i = 1; i <= P; i++) pw[i - 1] = ctrl_vibrato(s, i, pitch[i - 1]);
  SylCtrl &C = *ctrl_out;
  ctrl_sequential(s, anchors, z, A, C);
  if (C.status != SGB_OK) return C.status;
  for (int g = 0; g < C.nGC; g++) A.colmax[g] = ctrl_colmax(s, A, C, g);
  int kept = 0;
  for (int h = 1; h <= C.nHarmonics; h++) if (ctrl_rowkept(s, A, C, h)) A.rowmap[kept++] = h;
  C.rows_kept = kept;
  ctrl_sizes(A, C, 512);
  if (C.amp_elems > amp_cap) return SGB_ERR_INVALID;
  for (int e = 0; e < C.nEpochs; e++) {
    int rows = C.ep_rows[e];
    for (int g = C.ep_start[e] - 1; g < C.ep_end[e]; g++)
      for (int j = 1; j <= rows; j++)
        amp[C.ep_amp_off[e] + (int64_t)(g - (C.ep_start[e] - 1)) * rows + (j - 1)] = ampl_exact(s, A, C, e, j, g);
  }
  return SGB_OK;
}

int hs_sizeof_ctrl() { return (int)sizeof(SylCtrl); }
int hs_sizeof_syllable() { return (int)sizeof(sgb_syllable); }
}
