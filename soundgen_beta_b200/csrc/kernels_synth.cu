// K1: glottal-cycle additive synthesis -- the harmonic loop of generateHarmonics
// (R/source.R:389-419): waveform_epoch[k] = sum_rows sin(2*pi*integr[k]*times_f0) * am_upsampled[k].
//
// One CTA per (syllable, epoch, 512-sample tile).  Per sample, once:
//   * phase in FP64: integr = (phi_i + closed-form sum of the cubic piece of
//     pitch_upsampled) / samplingRate, re-anchored at every spline knot (one per
//     glottal cycle), so no error accumulates along the syllable;
//   * the stretched amplitude coordinate of approx() (source.R:403-405) -> (cycle, weight).
// Per (sample, row): rows are integer multiples j of theta' = 2*pi*integr/(nSubharm+1), so
// sum_j a_j sin(j theta') is evaluated with a blocked Clenshaw recurrence in Reinsch's
// stable form (4 FP32 FMAs per partial-sample, no sincos in the loop); every block of
// SYNTH_KBLOCK rows gets its base rotation e^{i j0 theta'} from a per-sample rotator whose
// step e^{i K theta'} is reduced in FP64 (no sincos in the loop).  The amplitude columns of
// the next row block are prefetched into registers while the current block computes.
#include "engine.cuh"

#define KPAD (SYNTH_KBLOCK + 1)

#ifndef SYNTH_MIN_CTAS
#define SYNTH_MIN_CTAS 5
#endif
#define SYNTH_NI_FAST 4      // cycles per tile handled by the double-buffered (prefetching) path
#define SYNTH_TAB 32         // spline pieces / cycle starts cached per tile
#define SYNTH_SB (2 * SYNTH_KBLOCK)   // rows per super-block: two Clenshaw blocks run interleaved
#define SBPAD (2 * KPAD)

struct SynthFix { double x; float w; int gi; int k; int pad; };   // a sample whose pair straddles a cycle

__global__ void __launch_bounds__(SYNTH_THREADS, SYNTH_MIN_CTAS)
k_synth(const SynthTile *__restrict__ tiles, const sgb_syllable *__restrict__ syl,
        const SylCtrl *__restrict__ ctrl, const SylLayout *__restrict__ lay, Pools P,
        const float2 *__restrict__ amp, float *__restrict__ wave, int *__restrict__ epmax) {
  // [cycle][half][row] {Y_g, Y_g, dY, dY}: one LDS.128 yields both packed operands of a pair
  __shared__ float4 sA[2 * SYNTH_NI_FAST * SBPAD];
  float4 *sBig = sA;   // the single-buffered path for up to SYNTH_NI_CAP cycles reuses the same storage
  static_assert(SYNTH_NI_CAP <= 2 * SYNTH_NI_FAST, "sBig aliases sA");
  __shared__ int sh_rng[4];
  __shared__ int t_gc[SYNTH_TAB + 2];
  __shared__ double t_rcp[SYNTH_TAB + 1];
  __shared__ double t_kt[SYNTH_TAB + 2], t_phi[SYNTH_TAB + 1], t_py[SYNTH_TAB + 1], t_sb[SYNTH_TAB + 1],
      t_sc[SYNTH_TAB + 1], t_sd[SYNTH_TAB + 1];
  __shared__ SynthFix fixl[SYNTH_TAB];
  __shared__ int n_fix;
  __shared__ float fix_red[SYNTH_THREADS / 32];

  const SynthTile T = tiles[blockIdx.x];
  const int s = T.syl, e = T.epoch, k0 = T.k0;
  const SylCtrl &C = ctrl[s];
  const double sr = syl[s].samplingRate;
  const int64_t o = P.gc_off[s];
  const int32_t *__restrict__ gcup = P.gcup + o;
  const double *__restrict__ kt = P.kt + o;
  const int G = C.nGC;
  const int g_first = C.ep_start[e] - 1;       // first gc of the epoch (0-based)
  const int g_lastStart = C.ep_end[e] - 1;     // last gc of the epoch
  const int nsub = C.vf_active ? C.ep_nsub[e] : 0;
  const int J = C.ep_rows[e];
  const int x_first_i = gcup[g_first];
  const double x_first = (double)x_first_i, x_last = (double)gcup[g_lastStart];
  const int Ne = gcup[C.ep_end[e]] - x_first_i + 1;
  const double by = (x_last - x_first) / (double)(Ne - 1);
  const int nknots_e = g_lastStart - g_first + 1;   // knots of approx(): gc starts of the epoch
  const float2 *__restrict__ ampE = amp + lay[s].amp_off + C.ep_amp_off[e];   // {Y_g, Y_{g+1} - Y_g}
  const double inv_sr_np1 = 1.0 / (sr * (double)(nsub + 1));
  const int klast_tile = min(k0 + SYNTH_TILE, Ne) - 1;

  // ---- tile tables: which cycles / spline pieces the tile touches (two threads search) ----
  if (threadIdx.x == 0) n_fix = 0;
  if (threadIdx.x < 2) {
    int k = threadIdx.x ? klast_tile : k0;
    double v = (k >= Ne - 1) ? x_last : (x_first + (double)k * by);
    int lo = 0, hi = nknots_e - 1;
    while (hi > lo + 1) {
      int mid = (lo + hi) >> 1;
      if (v < (double)gcup[g_first + mid]) hi = mid; else lo = mid;
    }
    sh_rng[threadIdx.x] = lo;
  } else if (threadIdx.x >= 32 && threadIdx.x < 34) {
    int k = (threadIdx.x & 1) ? klast_tile : k0;
    double u = (double)(x_first_i + k);
    int a = 0, b = G;
    while (b > a + 1) {
      int mid = (a + b) >> 1;
      if (u < kt[mid]) b = mid; else a = mid;
    }
    sh_rng[2 + (threadIdx.x & 1)] = a;
  }
  __syncthreads();
  const int gi_lo = sh_rng[0], n_int = sh_rng[1] - sh_rng[0] + 1;
  const int a_lo = sh_rng[2], n_pc = sh_rng[3] - sh_rng[2] + 1;
  const bool tab_ok = (n_int <= SYNTH_TAB) && (n_pc <= SYNTH_TAB);
  if (tab_ok) {
    for (int i = threadIdx.x; i <= n_int; i += SYNTH_THREADS) {
      int g0v = gcup[g_first + gi_lo + i];
      t_gc[i] = g0v;
      if (i < n_int) t_rcp[i] = 1.0 / (double)(gcup[g_first + gi_lo + i + 1] - g0v);
    }
    for (int i = threadIdx.x; i < n_pc; i += SYNTH_THREADS) {
      int a = a_lo + i;
      t_kt[i] = kt[a]; t_phi[i] = P.phi[o + a]; t_py[i] = P.ppg[o + a];
      t_sb[i] = P.sb[o + a]; t_sc[i] = P.sc[o + a]; t_sd[i] = P.sd[o + a];
    }
    if (threadIdx.x == 0) t_kt[n_pc] = (a_lo + n_pc < G) ? kt[a_lo + n_pc] : 1.0e300;
  }
  __syncthreads();

  // ---- per-sample set-up (4 consecutive samples per thread = 2 packed pairs) ----
  float w[SYNTH_SPT], delta[SYNTH_SPT], sigma[SYNTH_SPT], sint[SYNTH_SPT], rotc[SYNTH_SPT], rots[SYNTH_SPT];
  int gi[SYNTH_SPT];
  double xph[SYNTH_SPT];
  const int kbase = k0 + SYNTH_SPT * threadIdx.x;
#pragma unroll
  for (int i = 0; i < SYNTH_SPT; i++) {
    int k = kbase + i;
    if (k >= Ne) k = Ne - 1;
    // amplitude coordinate: seq(x_first, x_last, length.out = Ne)[k]
    double v = (k >= Ne - 1) ? x_last : (x_first + (double)k * by);
    double u = (double)(x_first_i + k);
    int lo, a;
    double xg, rcp, M, f_phi, f_py, f_sb, f_sc, f_sd;
    if (tab_ok) {
      lo = 0;
      while (lo < n_int - 1 && v >= (double)t_gc[lo + 1]) lo++;
      xg = (double)t_gc[lo]; rcp = t_rcp[lo];
      a = 0;
      while (a < n_pc - 1 && u >= t_kt[a + 1]) a++;
      M = u - t_kt[a];
      f_phi = t_phi[a]; f_py = t_py[a]; f_sb = t_sb[a]; f_sc = t_sc[a]; f_sd = t_sd[a];
    } else {
      int l2 = 0, hi = nknots_e - 1;
      while (hi > l2 + 1) {
        int mid = (l2 + hi) >> 1;
        if (v < (double)gcup[g_first + mid]) hi = mid; else l2 = mid;
      }
      xg = (double)gcup[g_first + l2]; rcp = 1.0 / ((double)gcup[g_first + l2 + 1] - xg);
      lo = l2 - gi_lo;
      int a2 = 0, b = G;
      while (b > a2 + 1) {
        int mid = (a2 + b) >> 1;
        if (u < kt[mid]) b = mid; else a2 = mid;
      }
      M = u - kt[a2];
      f_phi = P.phi[o + a2]; f_py = P.ppg[o + a2]; f_sb = P.sb[o + a2]; f_sc = P.sc[o + a2]; f_sd = P.sd[o + a2];
    }
    w[i] = (float)((v - xg) * rcp);
    gi[i] = lo;
    // phase (cycles) of sample u of the syllable: closed-form sum of the cubic spline piece
    double s1 = M * (M + 1.0) * 0.5;
    double s2 = M * (M + 1.0) * (2.0 * M + 1.0) / 6.0;
    double s3 = s1 * s1;
    double x = (f_phi + f_py * (M + 1.0) + f_sb * s1 + f_sc * s2 + f_sd * s3) * inv_sr_np1;   // integr / (n+1)
    x -= floor(x);
    xph[i] = x;
    double q = rint(2.0 * x);
    float xr = (float)(x - 0.5 * q);                    // in [-0.25, 0.25]
    float sg = (((int)q) & 1) ? -1.0f : 1.0f;
    float sh, ch;
    sincospif(xr, &sh, &ch);
    sigma[i] = sg;
    delta[i] = -sg * 4.0f * sh * sh;                    // 2cos(theta) - 2 sigma
    sint[i] = sg * 2.0f * sh * ch;                      // sin(theta)
    // e^{i K theta}: rotation between the bases of consecutive row blocks (FP64-reduced)
    double xk = (double)SYNTH_KBLOCK * x;
    xk -= rint(xk);
    sincospif(2.0f * (float)xk, &rots[i], &rotc[i]);
  }
  const bool use_smem = (n_int <= SYNTH_NI_CAP);
  const bool fast = (n_int <= SYNTH_NI_FAST);
  // A pair takes the amplitude column of its FIRST sample.  If the second sample already lies in
  // the next cycle its sum is recomputed after the main loop (at most n_int - 1 samples per tile).
#pragma unroll
  for (int p = 0; p < 2; p++) {
    if (gi[2 * p + 1] != gi[2 * p] && kbase + 2 * p + 1 < Ne && use_smem) {
      int slot = atomicAdd(&n_fix, 1);
      if (slot < SYNTH_TAB) {
        SynthFix f; f.x = xph[2 * p + 1]; f.w = w[2 * p + 1]; f.gi = gi[2 * p + 1];
        f.k = SYNTH_SPT * threadIdx.x + 2 * p + 1; f.pad = 0;
        fixl[slot] = f;
      }
    }
  }
  const int gp0 = gi[0], gp1 = gi[2];
  const bool one_col = __all_sync(0xffffffffu, gp0 == gp1);

  float2 w2[2], delta2[2], sigma2[2];
  w2[0] = make_float2(w[0], w[1]); w2[1] = make_float2(w[2], w[3]);
  delta2[0] = make_float2(delta[0], delta[1]); delta2[1] = make_float2(delta[2], delta[3]);
  sigma2[0] = make_float2(sigma[0], sigma[1]); sigma2[1] = make_float2(sigma[2], sigma[3]);

  float acc[SYNTH_SPT], basec[SYNTH_SPT], bases[SYNTH_SPT];
#pragma unroll
  for (int i = 0; i < SYNTH_SPT; i++) { acc[i] = 0.0f; basec[i] = 1.0f; bases[i] = 0.0f; }

  // prefetch registers of the double-buffered path: up to 4 elements per thread per super-block
  float2 pf[4];
  auto pf_load = [&](int j0) {
    const int mk = min(SYNTH_SB, J - j0);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int idx = threadIdx.x + r * SYNTH_THREADS;
      if (idx < n_int * SYNTH_SB) {
        int ii = idx / SYNTH_SB, m = idx - ii * SYNTH_SB;
        pf[r] = (m < mk) ? ampE[(int64_t)(gi_lo + ii) * J + j0 + m] : make_float2(0.0f, 0.0f);
      }
    }
  };
  auto pf_store = [&](float4 *dst) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      int idx = threadIdx.x + r * SYNTH_THREADS;
      if (idx < n_int * SYNTH_SB) {
        int ii = idx / SYNTH_SB, m = idx - ii * SYNTH_SB;
        int half = m / SYNTH_KBLOCK, mm = m - half * SYNTH_KBLOCK;
        dst[ii * SBPAD + half * KPAD + mm] = make_float4(pf[r].x, pf[r].x, pf[r].y, pf[r].y);
      }
    }
  };
  if (fast) pf_load(0);

  int blk = 0;
  for (int j0 = 0; j0 < J; j0 += SYNTH_SB, blk++) {
    const int mk = min(SYNTH_SB, J - j0);
    const float4 *tab;
    if (fast) {
      float4 *dst = sA + (blk & 1) * (SYNTH_NI_FAST * SBPAD);
      pf_store(dst);
      __syncthreads();
      if (j0 + SYNTH_SB < J) pf_load(j0 + SYNTH_SB);   // in flight while this super-block computes
      tab = dst;
    } else if (use_smem) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < n_int * SYNTH_SB; idx += SYNTH_THREADS) {
        int ii = idx / SYNTH_SB, m = idx - ii * SYNTH_SB;
        float2 q = (m < mk) ? ampE[(int64_t)(gi_lo + ii) * J + j0 + m] : make_float2(0.0f, 0.0f);
        int half = m / SYNTH_KBLOCK, mm = m - half * SYNTH_KBLOCK;
        sBig[ii * SBPAD + half * KPAD + mm] = make_float4(q.x, q.x, q.y, q.y);
      }
      __syncthreads();
      tab = sBig;
    } else {
      tab = sBig;
    }
    // two Clenshaw blocks (rows j0+1..j0+64 and j0+65..j0+128) x two pairs = 4 independent chains
    float2 bb2[4], dd2[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { bb2[c] = make_float2(0.0f, 0.0f); dd2[c] = make_float2(0.0f, 0.0f); }
    const int mtop = min(mk, SYNTH_KBLOCK);
    if (use_smem && one_col) {   // every thread of the warp has both pairs in one cycle: 2 loads per row pair
      const float4 *r0 = tab + gp0 * SBPAD;
#pragma unroll 4
      for (int m = mtop - 1; m >= 0; m--) {
        const float4 a0 = r0[m], b0 = r0[KPAD + m];
#define STEPY(CH, PR, Q)                                                                   \
        {                                                                                  \
          float2 a_ = __ffma2_rn(w2[PR], make_float2(Q.z, Q.w), make_float2(Q.x, Q.y));     \
          dd2[CH] = __ffma2_rn(delta2[PR], bb2[CH], __ffma2_rn(sigma2[PR], dd2[CH], a_));   \
          bb2[CH] = __ffma2_rn(sigma2[PR], bb2[CH], dd2[CH]);                              \
        }
        STEPY(0, 0, a0) STEPY(1, 1, a0) STEPY(2, 0, b0) STEPY(3, 1, b0)
      }
    } else if (use_smem) {
      const float4 *r0 = tab + gp0 * SBPAD, *r1 = tab + gp1 * SBPAD;
#pragma unroll 4
      for (int m = mtop - 1; m >= 0; m--) {
        const float4 a0 = r0[m], a1 = r1[m], b0 = r0[KPAD + m], b1 = r1[KPAD + m];
#define STEPX(CH, PR, Q)                                                                   \
        {                                                                                  \
          float2 a_ = __ffma2_rn(w2[PR], make_float2(Q.z, Q.w), make_float2(Q.x, Q.y));     \
          dd2[CH] = __ffma2_rn(delta2[PR], bb2[CH], __ffma2_rn(sigma2[PR], dd2[CH], a_));   \
          bb2[CH] = __ffma2_rn(sigma2[PR], bb2[CH], dd2[CH]);                              \
        }
        STEPX(0, 0, a0) STEPX(1, 1, a1) STEPX(2, 0, b0) STEPX(3, 1, b1)
      }
    } else {   // very high pitch: more cycles per tile than fit in shared memory, read L2 directly
      for (int m = mtop - 1; m >= 0; m--) {
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          float ya[SYNTH_SPT], da[SYNTH_SPT];
#pragma unroll
          for (int i = 0; i < SYNTH_SPT; i++) {
            int row = j0 + hf * SYNTH_KBLOCK + m;
            ya[i] = 0.0f; da[i] = 0.0f;
            if (row < J) {
              float2 q = ampE[(int64_t)(gi_lo + gi[i]) * J + row];
              ya[i] = q.x; da[i] = q.y;
            }
          }
#pragma unroll
          for (int pr = 0; pr < 2; pr++) {
            int ch = 2 * hf + pr;
            float2 a_ = __ffma2_rn(w2[pr], make_float2(da[2 * pr], da[2 * pr + 1]), make_float2(ya[2 * pr], ya[2 * pr + 1]));
            dd2[ch] = __ffma2_rn(delta2[pr], bb2[ch], __ffma2_rn(sigma2[pr], dd2[ch], a_));
            bb2[ch] = __ffma2_rn(sigma2[pr], bb2[ch], dd2[ch]);
          }
        }
      }
    }
    // block sums: S = b1 sin(theta), C = b1 delta/2 + sigma d1; contribution Im(e^{i j0 theta} (C + iS))
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const float bbs[SYNTH_SPT] = {bb2[2 * hf].x, bb2[2 * hf].y, bb2[2 * hf + 1].x, bb2[2 * hf + 1].y};
      const float dds[SYNTH_SPT] = {dd2[2 * hf].x, dd2[2 * hf].y, dd2[2 * hf + 1].x, dd2[2 * hf + 1].y};
#pragma unroll
      for (int i = 0; i < SYNTH_SPT; i++) {
        float Ss = bbs[i] * sint[i];
        float Cs = fmaf(bbs[i], 0.5f * delta[i], sigma[i] * dds[i]);
        acc[i] += fmaf(bases[i], Cs, basec[i] * Ss);
        float nc = fmaf(basec[i], rotc[i], -bases[i] * rots[i]);
        float ns = fmaf(basec[i], rots[i], bases[i] * rotc[i]);
        basec[i] = nc; bases[i] = ns;
      }
    }
  }

  // ---- samples whose pair straddled a cycle boundary: direct cooperative evaluation ----
  __syncthreads();
  const int nfx = min(n_fix, SYNTH_TAB);
  for (int f = 0; f < nfx; f++) {
    const SynthFix F = fixl[f];
    const float2 *col = ampE + (int64_t)(gi_lo + F.gi) * J;
    float part = 0.0f;
    for (int j = 1 + threadIdx.x; j <= J; j += SYNTH_THREADS) {
      float2 q = col[j - 1];
      float a_ = fmaf(F.w, q.y, q.x);
      double xj = (double)j * F.x;
      xj -= rint(xj);
      part = fmaf(a_, sinpif(2.0f * (float)xj), part);
    }
    for (int of = 16; of > 0; of >>= 1) part += __shfl_xor_sync(0xffffffffu, part, of);
    if ((threadIdx.x & 31) == 0) fix_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (F.k / SYNTH_SPT == (int)threadIdx.x) {
      float tot = 0.0f;
      for (int i = 0; i < SYNTH_THREADS / 32; i++) tot += fix_red[i];
      const int slot = F.k % SYNTH_SPT;
#pragma unroll
      for (int i = 0; i < SYNTH_SPT; i++) if (i == slot) acc[i] = tot;
    }
    __syncthreads();
  }

  // max |w| of the epoch (tolerance of the zero-crossing searches in K6)
  {
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) if (kbase + i < Ne) m = fmaxf(m, fabsf(acc[i]));
    for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));
    if ((threadIdx.x & 31) == 0) atomicMax(&epmax[(int64_t)s * SGB_MAX_EPOCHS + e], float_to_ordered(m));
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
                  const SylLayout *lay, const Pools &P, const float2 *amp, float *wave, int *epmax,
                  cudaStream_t st) {
  if (n_tiles <= 0) return;
  k_synth<<<n_tiles, SYNTH_THREADS, 0, st>>>(tiles, syl, ctrl, lay, P, amp, wave, epmax);
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
