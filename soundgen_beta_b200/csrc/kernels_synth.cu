// K1: glottal-cycle additive synthesis -- the harmonic loop of generateHarmonics
// (R/source.R:389-419): waveform_epoch[k] = sum_rows sin(2*pi*integr[k]*times_f0) * am_upsampled[k].
//
// One CTA per (syllable, epoch, 512-sample tile); its four warps are AUTONOMOUS: each owns 128
// consecutive samples (4 per lane = 2 packed pairs), stages its own amplitude rows and never meets
// a CTA barrier, so a warp that waits (set-up, copies) never holds up another warp's FMA stream.
// Per sample, once:
//   * phase in FP64: integr = (closed-form sum of the cubic piece of pitch_upsampled) / samplingRate,
//     a quartic in the offset from the spline knot (coefficients from K0), re-anchored at every knot
//     (one per glottal cycle), so no error accumulates along the syllable;
//   * the stretched amplitude coordinate of approx() (source.R:403-405) -> (cycle, weight).
// Per (sample, row): rows are integer multiples j of theta' = 2*pi*integr/(nSubharm+1), so
// sum_j a_j sin(j theta') is evaluated with a blocked Clenshaw recurrence in Reinsch's stable form
// (4 packed FP32 FMAs per pair of partial-samples, no sincos in the loop); every block of K rows
// gets its base rotation e^{i j0 theta'} from a per-sample rotator whose step e^{i K theta'} is
// reduced in FP64.  The {Y, Y, dY, dY} amplitude rows of the next block travel global -> shared
// memory with cp.async (16 B per lane, L1-allocating: the four warps of a tile read the same lines)
// into a per-warp double buffer while the current block computes; one broadcast LDS.128 per row
// feeds both packed operands of both pairs.  K = 128 / (cycles the warp touches, rounded up to a
// power of two) so that a stage always fits SYNTH_STAGE entries.
#include "engine.cuh"
#include <cstdlib>

#ifndef SYNTH_MIN_CTAS
#define SYNTH_MIN_CTAS 6
#endif
#define SYNTH_WARPS (SYNTH_THREADS / 32)
#define SYNTH_WSAMP (32 * SYNTH_SPT)      // samples per warp
#define SYNTH_KFALL 64                    // block length of the direct (no staging) path
#define FULLMASK 0xffffffffu

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Row recurrences for one packed pair (two samples).  The FMA pipe is limited by register-file reads
// (measured, scripts/micro/rf_model.cu: an FFMA2 with three fresh register pairs takes 3.1 cycles, with
// two 2.3, an FADD2 2.1), so each warp picks the cheapest form that is safe for ALL its lanes:
//   mode 3  standard Clenshaw  b <- c2 b1 - b2 + a            (3 ops) every lane at least 0.03 cycles
//           (11 degrees) away from theta = 0 and pi, where 2 cos(theta) carries the angle accurately;
//   mode 1/2 Reinsch, sigma = +1 / -1 for the whole warp       (4 ops, two of them adds): every lane
//           within 3/8 cycle of the pole the form is built around;
//   mode 0  Reinsch with a per-lane sigma                      (4 FMAs) anything else (a warp that spans
//           more than a quarter cycle of phase: high pitch).
#define SYNTH_LERP(CH, Q) __ffma2_rn(w2[CH], make_float2(Q.z, Q.w), make_float2(Q.x, Q.y))
#define SYNTH_STEP0(CH, Q)                                                                        \
  {                                                                                               \
    const float2 a_ = SYNTH_LERP(CH, Q);                                                          \
    s1[CH] = __ffma2_rn(k0[CH], s0[CH], __ffma2_rn(k1[CH], s1[CH], a_));                           \
    s0[CH] = __ffma2_rn(k1[CH], s0[CH], s1[CH]);                                                  \
  }
#define SYNTH_STEP1(CH, Q)                                                                        \
  {                                                                                               \
    const float2 a_ = SYNTH_LERP(CH, Q);                                                          \
    s1[CH] = __ffma2_rn(k0[CH], s0[CH], __fadd2_rn(s1[CH], a_));                                   \
    s0[CH] = __fadd2_rn(s0[CH], s1[CH]);                                                          \
  }
#define SYNTH_STEP2(CH, Q)                                                                        \
  {                                                                                               \
    const float2 a_ = SYNTH_LERP(CH, Q);                                                          \
    s1[CH] = __ffma2_rn(k0[CH], s0[CH], __fadd2_rn(a_, make_float2(-s1[CH].x, -s1[CH].y)));        \
    s0[CH] = __fadd2_rn(s1[CH], make_float2(-s0[CH].x, -s0[CH].y));                               \
  }
#define SYNTH_STEP3(CH, Q)                                                                        \
  {                                                                                               \
    const float2 a_ = SYNTH_LERP(CH, Q);                                                          \
    const float2 nb = __fadd2_rn(__ffma2_rn(k0[CH], s0[CH], a_), make_float2(-s1[CH].x, -s1[CH].y)); \
    s1[CH] = s0[CH]; s0[CH] = nb;                                                                 \
  }
// one block of mk rows for both pairs of a lane; r0 / r1: the pairs' amplitude columns in the stage
#define SYNTH_ROWS(STEP, UNR)                                                                     \
  if (one_col) {                                                                                  \
    _Pragma(UNR) for (int m = mk - 1; m >= 0; m--) { const float4 q = r0[m]; STEP(0, q) STEP(1, q) } \
  } else {                                                                                        \
    _Pragma(UNR) for (int m = mk - 1; m >= 0; m--) { const float4 q0 = r0[m], q1 = r1[m]; STEP(0, q0) STEP(1, q1) } \
  }

__global__ void __launch_bounds__(SYNTH_THREADS, SYNTH_MIN_CTAS)
k_synth(const SynthTile *__restrict__ tiles, const sgb_syllable *__restrict__ syl,
        const SylCtrl *__restrict__ ctrl, const SylLayout *__restrict__ lay, Pools P,
        const float4 *__restrict__ amp, float *__restrict__ wave, int *__restrict__ epmax, int force_mode) {
  __shared__ float4 sStage[SYNTH_WARPS][2][SYNTH_STAGE];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SynthTile T = tiles[blockIdx.x];
  const int s = T.syl, e = T.epoch;
  const SylCtrl &C = ctrl[s];
  const int64_t o = P.gc_off[s];
  const int32_t *__restrict__ gcup = P.gcup + o;
  const double *__restrict__ pcs = P.pc + SYNTH_PC * o;
  const int G = C.nGC;
  const int g_first = C.ep_start[e] - 1;       // first gc of the epoch (0-based)
  const int g_lastStart = C.ep_end[e] - 1;     // last gc of the epoch
  const int x_first_i = gcup[g_first];
  const int Ne = gcup[C.ep_end[e]] - x_first_i + 1;
  const int kw = T.k0 + SYNTH_WSAMP * wid;     // the warp's first sample
  if (kw >= Ne) return;                        // no CTA-wide barrier below
  const int nsub = C.vf_active ? C.ep_nsub[e] : 0;
  const int J = C.ep_rows[e];
  const double x_first = (double)x_first_i, x_last = (double)gcup[g_lastStart];
  const double by = (x_last - x_first) / (double)(Ne - 1);
  const int nknots_e = g_lastStart - g_first + 1;   // knots of approx(): gc starts of the epoch
  const float4 *__restrict__ ampE = amp + lay[s].amp_off + C.ep_amp_off[e];   // {Y_g, Y_g, dY, dY}
  const double inv_sr_np1 = 1.0 / (syl[s].samplingRate * (double)(nsub + 1));

  // ---- where the warp starts: amplitude interval and spline piece (forward scans from the tile's) ----
  int lo_w = T.gi_lo, a_w = T.a_lo;
  {
    const double v = (kw >= Ne - 1) ? x_last : (x_first + (double)kw * by);
    while (lo_w < nknots_e - 2 && v >= (double)gcup[g_first + lo_w + 1]) lo_w++;
    const double u = (double)(x_first_i + kw);
    while (a_w < G - 1 && u >= pcs[SYNTH_PC * (a_w + 1)]) a_w++;
  }

  // ---- per-sample set-up (4 consecutive samples per lane = 2 packed pairs) ----
  float w[SYNTH_SPT];
  int gi[SYNTH_SPT];
  double xph[SYNTH_SPT];
  const int kbase = kw + SYNTH_SPT * lane;
  {
    int lo = lo_w, a = a_w;
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) {
      int k = kbase + i;
      if (k >= Ne) k = Ne - 1;
      // amplitude coordinate: seq(x_first, x_last, length.out = Ne)[k]
      const double v = (k >= Ne - 1) ? x_last : (x_first + (double)k * by);
      const double u = (double)(x_first_i + k);
      while (lo < nknots_e - 2 && v >= (double)gcup[g_first + lo + 1]) lo++;
      while (a < G - 1 && u >= pcs[SYNTH_PC * (a + 1)]) a++;
      const int xg = gcup[g_first + lo], xn = gcup[g_first + lo + 1];
      w[i] = (float)(v - (double)xg) * __frcp_rn((float)(xn - xg));
      gi[i] = lo - lo_w;
      // phase (cycles) of sample u of the syllable: quartic in the offset from the knot
      const double *pp = pcs + SYNTH_PC * a;
      const double M = u - pp[0];
      const double ph = fma(fma(fma(fma(pp[5], M, pp[4]), M, pp[3]), M, pp[2]), M, pp[1]);
      double x = ph * inv_sr_np1;               // integr / (nSubharm + 1)
      x -= floor(x);
      xph[i] = x;
    }
  }
  const int ncolw = __shfl_sync(FULLMASK, gi[SYNTH_SPT - 1], 31) + 1;   // cycles the warp touches
  const bool staged = (ncolw <= 16);
  int kshift = 7;                                                      // K = 128 >> ceil(log2(ncolw))
  if (staged) { for (int cap = 1; cap < ncolw; cap <<= 1) kshift--; }
  else kshift = 31 - __clz(SYNTH_KFALL);
  if ((1 << kshift) > SYNTH_KMAX) kshift = 31 - __clz(SYNTH_KMAX);
  const int K = 1 << kshift;

  // which recurrence is safe for every lane of the warp (distances in cycles from theta = 0 / pi)
  int mode;
  {
    bool okA = true, okP = true, okM = true;
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) {
      const float xf = (float)xph[i];
      const float d0 = fminf(xf, 1.0f - xf), dpi = fabsf(xf - 0.5f);
      okA = okA && (d0 >= 0.03f) && (dpi >= 0.03f);
      okP = okP && (d0 <= 0.375f);
      okM = okM && (dpi <= 0.375f);
    }
    okA = __all_sync(FULLMASK, okA); okP = __all_sync(FULLMASK, okP); okM = __all_sync(FULLMASK, okM);
    mode = okA ? 3 : (okP ? 1 : (okM ? 2 : 0));
    if (force_mode >= 0) mode = force_mode;   // test hook (SGB_SYNTH_MODE): 0 is valid for every warp
  }
  // k0: delta = 2cos(theta) - 2 sigma (Reinsch) or 2cos(theta) (standard); k1: sigma; ec/es: cos, sin(theta)
  float2 w2[2], k0[2], k1[2], ec2[2], es2[2], rotc2[2], rots2[2];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    float kk[2], sg[2], co[2], si[2], rc[2], rs[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const double x = xph[2 * p + h];
      const double q = rint(2.0 * x);
      const float xr = (float)(x - 0.5 * q);              // in [-0.25, 0.25]
      float sgn = (((int)q) & 1) ? -1.0f : 1.0f;          // the nearer pole
      float sh, ch;
      sincospif(xr, &sh, &ch);
      si[h] = sgn * 2.0f * sh * ch;                       // sin(theta)
      co[h] = sgn * fmaf(-2.0f * sh, sh, 1.0f);           // cos(theta)
      float dl = -sgn * 4.0f * sh * sh;                   // 2cos(theta) - 2 sgn, accurate near the pole sgn
      if (mode == 1 && sgn < 0.0f) { dl = -4.0f * ch * ch; sgn = 1.0f; }   // 2cos(theta) - 2: form built around theta = 0
      if (mode == 2 && sgn > 0.0f) { dl = 4.0f * ch * ch; sgn = -1.0f; }   // 2cos(theta) + 2: ... around theta = pi
      kk[h] = (mode == 3) ? 2.0f * co[h] : dl;
      sg[h] = sgn;
      double xk = (double)K * x;                          // e^{i K theta}: base rotation between row blocks
      xk -= rint(xk);
      sincospif(2.0f * (float)xk, &rs[h], &rc[h]);
    }
    w2[p] = make_float2(w[2 * p], w[2 * p + 1]);
    k0[p] = make_float2(kk[0], kk[1]); k1[p] = make_float2(sg[0], sg[1]);
    ec2[p] = make_float2(co[0], co[1]); es2[p] = make_float2(si[0], si[1]);
    rotc2[p] = make_float2(rc[0], rc[1]); rots2[p] = make_float2(rs[0], rs[1]);
  }

  // ---- a pair takes the amplitude column of its FIRST sample; if the second sample already lies in
  //      the next cycle its sum is evaluated directly, by the whole warp, before the main loop ----
  float fixv[2] = {0.0f, 0.0f};
  bool need[2];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    need[p] = staged && (gi[2 * p + 1] != gi[2 * p]) && (kbase + 2 * p + 1 < Ne);
    unsigned msk = __ballot_sync(FULLMASK, need[p]);
    while (msk) {
      const int src = __ffs(msk) - 1;
      msk &= msk - 1;
      const double xs = __shfl_sync(FULLMASK, xph[2 * p + 1], src);
      const float ws = __shfl_sync(FULLMASK, w[2 * p + 1], src);
      const int gs = __shfl_sync(FULLMASK, gi[2 * p + 1], src);
      const float4 *col = ampE + (int64_t)(lo_w + gs) * J;
      // every lane sums a contiguous run of rows with a float rotator started from an FP64-reduced phase
      const int chunk = (J + 31) >> 5;
      const int ja = lane * chunk + 1, jb = min(J, ja + chunk - 1);
      float part = 0.0f;
      if (ja <= jb) {
        float s1, c1, sj, cj;
        sincospif(2.0f * (float)(xs - rint(xs)), &s1, &c1);     // e^{i theta}
        double xa = (double)ja * xs;
        xa -= rint(xa);
        sincospif(2.0f * (float)xa, &sj, &cj);                  // e^{i ja theta}
        for (int j = ja; j <= jb; j++) {
          const float4 q = col[j - 1];
          part = fmaf(fmaf(ws, q.z, q.x), sj, part);
          const float cn = fmaf(cj, c1, -sj * s1), sn = fmaf(sj, c1, cj * s1);
          cj = cn; sj = sn;
        }
      }
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) part += __shfl_xor_sync(FULLMASK, part, of);
      if (lane == src) fixv[p] = part;
    }
  }

  float2 acc2[2], basec2[2], bases2[2];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    acc2[p] = make_float2(0.0f, 0.0f); basec2[p] = make_float2(1.0f, 1.0f); bases2[p] = make_float2(0.0f, 0.0f);
  }
  // block sums over local rows r = 1..mk: S = sum a_r sin(r theta) = b1 sin(theta); C = sum a_r cos(r theta)
  //   = b1 cos(theta) - b2 (standard) = b1 (cos(theta) - sigma) + sigma d1 (Reinsch, d1 = b1 - sigma b2);
  // contribution Im(e^{i j0 theta} (C + iS)); then the base advances by e^{i K theta}
#define SYNTH_EPILOGUE()                                                                              \
  _Pragma("unroll") for (int p = 0; p < 2; p++) {                                                     \
    const float2 Ss = __fmul2_rn(s0[p], es2[p]);                                                      \
    float2 Cs;                                                                                        \
    if (mode == 3) Cs = __ffma2_rn(s0[p], ec2[p], make_float2(-s1[p].x, -s1[p].y));                   \
    else Cs = __ffma2_rn(s0[p], __fmul2_rn(k0[p], make_float2(0.5f, 0.5f)), __fmul2_rn(k1[p], s1[p])); \
    acc2[p] = __ffma2_rn(bases2[p], Cs, __ffma2_rn(basec2[p], Ss, acc2[p]));                          \
    const float2 nb = make_float2(-bases2[p].x, -bases2[p].y);                                        \
    const float2 nc = __ffma2_rn(basec2[p], rotc2[p], __fmul2_rn(nb, rots2[p]));                      \
    const float2 ns = __ffma2_rn(basec2[p], rots2[p], __fmul2_rn(bases2[p], rotc2[p]));               \
    basec2[p] = nc; bases2[p] = ns;                                                                   \
  }

  if (staged) {
    float4(*stage)[SYNTH_STAGE] = sStage[wid];
    auto issue = [&](int buf, int j0) {
      float4 *dst = stage[buf];
#pragma unroll
      for (int r = 0; r < SYNTH_STAGE / 32; r++) {
        const int idx = lane + 32 * r;
        const int c = idx >> kshift, m = idx & (K - 1);
        if (c < ncolw && j0 + m < J) cp_async16(dst + idx, ampE + (int64_t)(lo_w + c) * J + j0 + m);
      }
      cp_async_commit();
    };
    issue(0, 0);
    const int c0 = gi[0] << kshift, c1 = gi[2] << kshift;
    const bool one_col = __all_sync(FULLMASK, c0 == c1);
    int blk = 0;
    for (int j0 = 0; j0 < J; j0 += K, blk++) {
      const int mk = min(K, J - j0);
      if (j0 + K < J) { issue((blk + 1) & 1, j0 + K); cp_async_wait<1>(); }   // next block in flight
      else cp_async_wait<0>();
      __syncwarp();
      const float4 *tab = stage[blk & 1];
      const float4 *r0 = tab + c0, *r1 = tab + c1;
      float2 s0[2], s1[2];
      s0[0] = s0[1] = s1[0] = s1[1] = make_float2(0.0f, 0.0f);
      if (mode == 3) { SYNTH_ROWS(SYNTH_STEP3, "unroll 8") }
      else if (mode == 1) { SYNTH_ROWS(SYNTH_STEP1, "unroll 8") }
      else if (mode == 2) { SYNTH_ROWS(SYNTH_STEP2, "unroll 8") }
      else { SYNTH_ROWS(SYNTH_STEP0, "unroll 4") }
      SYNTH_EPILOGUE()
      __syncwarp();      // the stage is rewritten by the copy issued in the next iteration
    }
  } else {
    // very high pitch (more than 16 cycles in 128 samples): every sample reads its own column from L1/L2
    const float4 *colp[SYNTH_SPT];
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) colp[i] = ampE + (int64_t)(lo_w + gi[i]) * J;
    for (int j0 = 0; j0 < J; j0 += K) {
      const int mk = min(K, J - j0);
      float2 s0[2], s1[2];
      s0[0] = s0[1] = s1[0] = s1[1] = make_float2(0.0f, 0.0f);
      for (int m = mk - 1; m >= 0; m--) {
        const float4 q0 = colp[0][j0 + m], q1 = colp[1][j0 + m], q2 = colp[2][j0 + m], q3 = colp[3][j0 + m];
        const float4 qa = make_float4(q0.x, q1.x, q0.z, q1.z), qb = make_float4(q2.x, q3.x, q2.z, q3.z);
        if (mode == 3) { SYNTH_STEP3(0, qa) SYNTH_STEP3(1, qb) }
        else { SYNTH_STEP0(0, qa) SYNTH_STEP0(1, qb) }
      }
      SYNTH_EPILOGUE()
    }
  }

  float acc[SYNTH_SPT] = {acc2[0].x, need[0] ? fixv[0] : acc2[0].y, acc2[1].x, need[1] ? fixv[1] : acc2[1].y};
  // max |w| of the epoch (tolerance of the zero-crossing searches in K6)
  {
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) if (kbase + i < Ne) m = fmaxf(m, fabsf(acc[i]));
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(FULLMASK, m, of));
    if (lane == 0) atomicMax(&epmax[(int64_t)s * SGB_MAX_EPOCHS + e], float_to_ordered(m));
  }
  float *out = wave + lay[s].wave_off + C.ep_wave_off[e];
  if (kbase + SYNTH_SPT <= Ne) {
    *reinterpret_cast<float4 *>(out + kbase) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else {
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) if (kbase + i < Ne) out[kbase + i] = acc[i];
  }
}

void launch_synth(const SynthTile *tiles, int n_tiles, const sgb_syllable *syl, const SylCtrl *ctrl,
                  const SylLayout *lay, const Pools &P, const float4 *amp, float *wave, int *epmax,
                  cudaStream_t st) {
  if (n_tiles <= 0) return;
  static const int force_mode = [] { const char *e = getenv("SGB_SYNTH_MODE"); return e ? atoi(e) : -1; }();
  k_synth<<<n_tiles, SYNTH_THREADS, 0, st>>>(tiles, syl, ctrl, lay, P, amp, wave, epmax, force_mode == 0 ? 0 : -1);
}

// FP32 pipe peak: 8 independent FFMA2 dependency chains per thread.
__global__ void __launch_bounds__(256) k_fp32_peak(float2 *out, int iters) {
  float2 a[8];
  const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-9f, -1e-9f);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = make_float2(1.0f + i + threadIdx.x * 1e-3f, 2.0f + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], m, c);
  }
  float2 s = a[0];
#pragma unroll
  for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
  if (s.x == 123.456f) out[threadIdx.x] = s;   // never true: keeps the chains alive
}
void launch_fp32_peak(float2 *out, int iters, int blocks, int threads) {
  k_fp32_peak<<<blocks, threads>>>(out, iters);
}
