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
// SYNTH_KBLOCK rows is re-anchored with an exact FP64-reduced rotation e^{i j0 theta'}.
#include "engine.cuh"

#define KPAD (SYNTH_KBLOCK + 1)

__global__ void __launch_bounds__(SYNTH_THREADS)
k_synth(const SynthTile *__restrict__ tiles, const sgb_syllable *__restrict__ syl,
        const SylCtrl *__restrict__ ctrl, const SylLayout *__restrict__ lay, Pools P,
        const double *__restrict__ amp, float *__restrict__ wave) {
  __shared__ float2 sA[SYNTH_NI_CAP * KPAD];   // [interval][row] {Y_g, Y_{g+1} - Y_g}
  __shared__ int sh_gi[2];

  const SynthTile T = tiles[blockIdx.x];
  const int s = T.syl, e = T.epoch, k0 = T.k0;
  const SylCtrl &C = ctrl[s];
  const double sr = syl[s].samplingRate;
  const int64_t o = P.gc_off[s];
  const int32_t *__restrict__ gcup = P.gcup + o;
  const double *__restrict__ kt = P.kt + o;
  const double *__restrict__ py = P.ppg + o;
  const double *__restrict__ sb = P.sb + o;
  const double *__restrict__ sc = P.sc + o;
  const double *__restrict__ sd = P.sd + o;
  const double *__restrict__ phi = P.phi + o;
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
  const double *__restrict__ ampE = amp + lay[s].amp_off + C.ep_amp_off[e];
  const double inv_np1 = (double)(nsub + 1);

  // ---- per-sample set-up (4 consecutive samples per thread) ----
  float w[SYNTH_SPT], delta[SYNTH_SPT], sigma[SYNTH_SPT], sint[SYNTH_SPT];
  double xph[SYNTH_SPT];
  int gi[SYNTH_SPT];
  const int kbase = k0 + SYNTH_SPT * threadIdx.x;
#pragma unroll
  for (int i = 0; i < SYNTH_SPT; i++) {
    int k = kbase + i;
    if (k >= Ne) k = Ne - 1;
    // amplitude coordinate: seq(x_first, x_last, length.out = Ne)[k]
    double v = (k >= Ne - 1) ? x_last : (x_first + (double)k * by);
    // largest knot index q in [0, nknots_e-2] with gcup[g_first+q] <= v
    int lo = 0, hi = nknots_e - 1;
    while (hi > lo + 1) {
      int mid = (lo + hi) >> 1;
      if (v < (double)gcup[g_first + mid]) hi = mid; else lo = mid;
    }
    double xg = (double)gcup[g_first + lo], xn = (double)gcup[g_first + lo + 1];
    w[i] = (float)((v - xg) / (xn - xg));
    gi[i] = lo;
    // phase (cycles) of sample u = x_first + k of the syllable
    double u = (double)(x_first_i + k);
    int a = 0, b = G;
    while (b > a + 1) {
      int mid = (a + b) >> 1;
      if (u < kt[mid]) b = mid; else a = mid;
    }
    double M = u - kt[a];
    double s1 = M * (M + 1.0) * 0.5;
    double s2 = M * (M + 1.0) * (2.0 * M + 1.0) / 6.0;
    double s3 = s1 * s1;
    double integr = (phi[a] + py[a] * (M + 1.0) + sb[a] * s1 + sc[a] * s2 + sd[a] * s3) / sr;
    double x = integr / inv_np1;
    x -= floor(x);
    xph[i] = x;
    double q = rint(2.0 * x);
    float xr = (float)(x - 0.5 * q);                    // in [-0.25, 0.25]
    float sg = (((int)q) & 1) ? -1.0f : 1.0f;
    float sh = sinpif(xr);
    sigma[i] = sg;
    delta[i] = -sg * 4.0f * sh * sh;                    // 2cos(theta) - 2 sigma
    sint[i] = sg * sinpif(2.0f * xr);                   // sin(theta)
  }
  if (threadIdx.x == 0) sh_gi[0] = gi[0];
  {
    int klast = min(k0 + SYNTH_TILE, Ne) - 1;
    if (kbase <= klast && klast < kbase + SYNTH_SPT) sh_gi[1] = gi[klast - kbase];
  }
  __syncthreads();
  const int gi_lo = sh_gi[0], n_int = sh_gi[1] - sh_gi[0] + 1;
  const bool use_smem = (n_int <= SYNTH_NI_CAP);
#pragma unroll
  for (int i = 0; i < SYNTH_SPT; i++) gi[i] -= gi_lo;

  float acc[SYNTH_SPT];
#pragma unroll
  for (int i = 0; i < SYNTH_SPT; i++) acc[i] = 0.0f;

  for (int j0 = 0; j0 < J; j0 += SYNTH_KBLOCK) {
    const int mk = min(SYNTH_KBLOCK, J - j0);
    if (use_smem) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < n_int * mk; idx += SYNTH_THREADS) {
        int ii = idx / mk, m = idx - ii * mk;
        const double *col = ampE + (int64_t)(gi_lo + ii) * J + j0 + m;
        double y0 = col[0], y1 = col[J];
        sA[ii * KPAD + m] = make_float2((float)y0, (float)(y1 - y0));
      }
      __syncthreads();
    }
    float bb[SYNTH_SPT], dd[SYNTH_SPT];
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) { bb[i] = 0.0f; dd[i] = 0.0f; }
    if (use_smem) {
      for (int m = mk - 1; m >= 0; m--) {
#pragma unroll
        for (int i = 0; i < SYNTH_SPT; i++) {
          float2 a2 = sA[gi[i] * KPAD + m];
          float a = fmaf(w[i], a2.y, a2.x);
          dd[i] = fmaf(delta[i], bb[i], fmaf(sigma[i], dd[i], a));
          bb[i] = fmaf(sigma[i], bb[i], dd[i]);
        }
      }
    } else {   // very high pitch: more than SYNTH_NI_CAP cycles per tile, read L2 directly
      for (int m = mk - 1; m >= 0; m--) {
#pragma unroll
        for (int i = 0; i < SYNTH_SPT; i++) {
          const double *col = ampE + (int64_t)(gi_lo + gi[i]) * J + j0 + m;
          double y0 = col[0], y1 = col[J];
          float a = fmaf(w[i], (float)(y1 - y0), (float)y0);
          dd[i] = fmaf(delta[i], bb[i], fmaf(sigma[i], dd[i], a));
          bb[i] = fmaf(sigma[i], bb[i], dd[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < SYNTH_SPT; i++) {
      float Ss = bb[i] * sint[i];
      if (j0 == 0) {
        acc[i] += Ss;
      } else {
        float Cs = fmaf(bb[i], 0.5f * delta[i], sigma[i] * dd[i]);
        double xb = (double)j0 * xph[i];
        xb -= rint(xb);
        float sbv, cbv;
        sincospif(2.0f * (float)xb, &sbv, &cbv);
        acc[i] += fmaf(sbv, Cs, cbv * Ss);
      }
    }
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
                  const SylLayout *lay, const Pools &P, const double *amp, float *wave, cudaStream_t st) {
  if (n_tiles <= 0) return;
  k_synth<<<n_tiles, SYNTH_THREADS, 0, st>>>(tiles, syl, ctrl, lay, P, amp, wave);
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
