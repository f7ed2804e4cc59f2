// Sample-rate glue of the bout orchestrator (R/soundgen.R:700-849): noise
// normalisation (source.R:125-131), mixing with addVectors (utilities_math.R:500-526),
// global amplitude envelope, normalisation after the filter, AM trill, silences.
#include "engine.cuh"
#include "contour.cuh"

// breathingStrength = getSmoothContour(len, noiseAnchors, floor -120, ceiling 40, samplingRate): the fit (FMM
// spline or loess: a millisecond of FP64 on one thread) once per noise, one thread each, instead of once per CTA of
// k_noise_final with 255 threads waiting
__global__ void k_noise_tabs(const sgb_noise *__restrict__ noises, int n_noise, const double *__restrict__ anchors,
                             ContourTab *__restrict__ tabs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_noise) return;
  const sgb_noise N = noises[n];
  if (N.strength_pre_off < 0)
    contour_prepare(&tabs[n], anchors + 2 * N.anchor_off, N.anchor_n, N.len, N.samplingRate, true, -120.0, true, 40.0,
                    false, N.anchor_method);
}

// generateNoise tail (R/source.R:70-81,125-131):
// breathing / max(breathing) * 2^(contour/10), then fadeInOut.  grid (noises, chunks).
__global__ void __launch_bounds__(256)
k_noise_final(const sgb_noise *__restrict__ noises, const NoiseLayout *__restrict__ nl,
              const ContourTab *__restrict__ tabs, const double *__restrict__ pre,
              const int *__restrict__ maxpool, int max_base, const float *__restrict__ raw,
              float *__restrict__ fin) {
  const int n = blockIdx.x;
  const sgb_noise N = noises[n];
  const int L = N.len;
  const float mx = ordered_to_float(maxpool[max_base + n]);
  const double rcp_mx = 1.0 / (double)mx;     // x * (1 / max): within an ulp of x / max in double, stored as FP32
  const float *src = raw + nl[n].raw_off;
  float *dst = fin + nl[n].raw_off;
  int lf = (int)floor(N.attackLen * N.samplingRate / 1000.0);
  if (lf < 2) lf = 0;
  if (lf > L) lf = L;
  __shared__ ContourTab T;
  if (N.strength_pre_off < 0) {
    static_assert(sizeof(ContourTab) % 8 == 0, "ContourTab is copied as doubles");
    const double *g = reinterpret_cast<const double *>(&tabs[n]);
    double *d = reinterpret_cast<double *>(&T);
    for (int i = threadIdx.x; i < (int)(sizeof(ContourTab) / 8); i += blockDim.x) d[i] = g[i];
    __syncthreads();
  }
  for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < L; k += gridDim.y * blockDim.x) {
    double c = (N.strength_pre_off >= 0) ? pre[N.strength_pre_off + k] : contour_eval(&T, L, k);
    double v = (double)src[k] * rcp_mx * exp2(c / 10.0);
    if (lf > 0) {
      if (k < lf) v = v * r_seq_at(0.0, 1.0, lf, k);
      if (k >= L - lf) v = v * r_seq_at(0.0, 1.0, lf, (L - 1) - k);
    }
    dst[k] = (float)v;
  }
}

// sound = voiced (+ breathing noise added BEFORE filtering, soundgen.R:708-714), then the
// global amplitude envelope (soundgen.R:721-733).  The voiced syllables are already in
// place; this pass adds the noises in call order and applies the envelope.
// grid (chunks, bouts).
__global__ void __launch_bounds__(256)
k_sound_mix(const sgb_bout *__restrict__ bouts, const BoutLayout *__restrict__ bl,
            const sgb_noise *__restrict__ noises, const NoiseLayout *__restrict__ nl,
            const double *__restrict__ anchors, const float *__restrict__ noise_fin,
            float *__restrict__ sound) {
  const int b = blockIdx.x;
  const sgb_bout B = bouts[b];
  const BoutLayout L = bl[b];
  float *snd = sound + L.sound_off;
  const bool has_env = (B.aglobal_n > 0);
  bool any_noise = false;
  for (int n = B.noise_begin; n < B.noise_end; n++) if (noises[n].mix == 0) any_noise = true;
  if (!any_noise && !has_env) return;
  // amplEnvelope = getSmoothContour(amplAnchorsGlobal, len = length(sound), 0, -throwaway, samplingRate)
  __shared__ ContourTab T;
  if (has_env) {
    if (threadIdx.x == 0)
      contour_prepare(&T, anchors + 2 * B.aglobal_off, B.aglobal_n, L.sound_len, B.samplingRate, true, 0.0, true,
                      -B.throwaway, false, B.aglobal_method);
    __syncthreads();
  }
  for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < L.sound_len; k += gridDim.y * blockDim.x) {
    float v = snd[k];
    for (int n = B.noise_begin; n < B.noise_end; n++) {
      if (noises[n].mix != 0) continue;
      int64_t rel = (L.sound_off + k) - nl[n].dst_off;
      if (rel >= 0 && rel < noises[n].len) v = v + noise_fin[nl[n].raw_off + rel];
    }
    if (has_env) {
      double e = contour_eval(&T, L.sound_len, k);
      v = (float)((double)v * e);
    }
    snd[k] = v;
  }
}

// soundFiltered / max, + separately filtered noise (soundgen.R:807-818), AM trill
// (soundgen.R:821-833, getSigmoid utilities_math.R:639-653), written at the bout's place in
// the call's output (silences were zero-filled).  grid (chunks, bouts).
template <typename OutT>
__global__ void __launch_bounds__(256)
k_finalize(const sgb_bout *__restrict__ bouts, const BoutLayout *__restrict__ bl,
           const sgb_noise *__restrict__ noises, const NoiseLayout *__restrict__ nl,
           const float *__restrict__ sound, const float *__restrict__ filt,
           const float *__restrict__ noise_fin, const int *__restrict__ maxpool,
           OutT *__restrict__ out) {
  const int b = blockIdx.x;
  const sgb_bout B = bouts[b];
  const BoutLayout L = bl[b];
  const float *src = L.bypass ? (sound + L.sound_off) : (filt + L.filt_off);
  const double mx = L.bypass ? 1.0 : (double)ordered_to_float(maxpool[b]);
  const double rcp_mx = 1.0 / mx;
  OutT *dst = out + L.out_off;
  // AM pattern
  const bool am = B.amDep > 0.0;
  int nb = 1;
  double from = 0, to = 0, slope = 0, bmin = 0, bmax = 1;
  if (am) {
    from = -exp(-B.amShape * 1.0);
    to = exp(B.amShape * 1.0);
    slope = exp(fabs(B.amShape)) * 5.0;
    nb = (int)ceil(B.samplingRate / B.amFreq / 2.0);
    if (nb < 1) nb = 1;
    bmin = 1.0 / (1.0 + exp(-from * slope));
    bmax = 1.0 / (1.0 + exp(-to * slope));
  }
#pragma unroll 4
  for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < L.final_len; k += gridDim.y * blockDim.x) {
    double v = 0.0;
    int rel = k - L.final_shift;
    if (rel >= 0 && rel < L.filt_len) v = (double)src[rel] * rcp_mx;
    for (int n = B.noise_begin; n < B.noise_end; n++) {
      if (noises[n].mix != 1) continue;
      int64_t r2 = (L.out_off + k) - nl[n].dst_off;
      if (r2 >= 0 && r2 < noises[n].len) v = v + (double)noise_fin[nl[n].raw_off + r2];
    }
    if (am) {
      int ph = k % (2 * nb);
      int i = (ph < nb) ? ph : (2 * nb - 1 - ph);
      double a = r_seq_at(from, to, nb, i);
      double bb = 1.0 / (1.0 + exp(-a * slope));
      double sig = (bb - bmin) / (bmax - bmin);
      v = v * (1.0 - sig * B.amDep / 100.0);
    }
    dst[k] = (OutT)v;
  }
}

void launch_noise_final(const sgb_noise *noises, int n_noise, const NoiseLayout *nl, const double *anchors,
                        const double *pre, const int *maxpool, int max_base, const float *raw, float *fin,
                        void *tabs, int chunks, cudaStream_t st) {
  if (n_noise <= 0) return;
  k_noise_tabs<<<(n_noise + 31) / 32, 32, 0, st>>>(noises, n_noise, anchors, (ContourTab *)tabs);
  dim3 g(n_noise, chunks);
  k_noise_final<<<g, 256, 0, st>>>(noises, nl, (const ContourTab *)tabs, pre, maxpool, max_base, raw, fin);
}
size_t contour_tab_bytes() { return sizeof(ContourTab); }

void launch_sound_mix(const sgb_bout *bouts, int n_bouts, const BoutLayout *bl, const sgb_noise *noises,
                      const NoiseLayout *nl, const double *anchors, const float *noise_fin, float *sound,
                      int chunks, cudaStream_t st) {
  if (n_bouts <= 0) return;
  dim3 g(n_bouts, chunks);
  k_sound_mix<<<g, 256, 0, st>>>(bouts, bl, noises, nl, anchors, noise_fin, sound);
}

void launch_finalize(int f64, const sgb_bout *bouts, int n_bouts, const BoutLayout *bl,
                     const sgb_noise *noises, const NoiseLayout *nl, const float *sound, const float *filt,
                     const float *noise_fin, const int *maxpool, void *out, int chunks, cudaStream_t st) {
  if (n_bouts <= 0) return;
  dim3 g(n_bouts, chunks);
  if (f64) k_finalize<double><<<g, 256, 0, st>>>(bouts, bl, noises, nl, sound, filt, noise_fin, maxpool, (double *)out);
  else k_finalize<float><<<g, 256, 0, st>>>(bouts, bl, noises, nl, sound, filt, noise_fin, maxpool, (float *)out);
}
