// K2: fused STFT -> envelope multiply -> ISTFT overlap-add (R/soundgen.R:743-806 with
// seewave::stft / istft, seewave.r:7782-7818 and :3447-3487), and
// K5: noise spectrum -> ISTFT overlap-add (R/source.R:88-124).
//
// One CTA walks a run of consecutive frames of one sound.  Two real frames share one
// complex FFT (A + iB); the FFT is a hand-written shared-memory Stockham autosort with
// radix-4 / radix-2 register butterflies and a table-driven generic radix for the odd
// factors soundgen's window sizes need (3, 5, 19, 29, or any prime after the clamp at
// soundgen.R:743).  Frame inputs are staged by TMA bulk copies (cp.async.bulk +
// mbarrier) into a double buffer while the previous pair is transformed; the weighted
// overlap-add lives in a shared-memory ring and every output sample is written once.
#include "engine.cuh"
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ int g_stft_timeout = 0;   // set when a staged frame load never completed (checked by the host)
int stft_timeout_flag() { int v = 0; cudaMemcpyFromSymbol(&v, g_stft_timeout, sizeof v); return v; }
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long polls = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 " selp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!ok && ++polls > (1ull << 26)) {   // seconds: a staged load that never lands must fail loudly, not hang
      g_stft_timeout = 1;
      return false;
    }
  }
  return true;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// Input of the first forward pass: two real frames read straight from the TMA staging buffer, multiplied by
// the analysis window and packed as A + iB (no separate windowing pass).
struct WinIn {
  const float *a, *b, *w;
  int off;
  bool hasB;
  __device__ __forceinline__ WinIn operator+(int o) const { WinIn r = *this; r.off += o; return r; }
  __device__ __forceinline__ float2 operator[](int i) const {
    const int k = off + i;
    const float wv = w[k];
    return make_float2(a[k] * wv, hasB ? b[k] * wv : 0.0f);
  }
};
// Output of the last inverse pass: weighted overlap-add straight into the rings (frame A of a pair into
// ring A, frame B into ring B, so no two threads of a pass touch the same slot).
struct OlaOut {
  float *ra, *rb;
  const float *ws;
  int oA, oB, ring;
  bool hasB;
  __device__ __forceinline__ void add(int o, float2 v) const {
    const float wv = ws[o];
    int sa = oA + o;
    if (sa >= ring) sa -= ring;
    ra[sa] += v.x * wv;
    if (hasB) {
      int sb = oB + o;
      if (sb >= ring) sb -= ring;
      rb[sb] += v.y * wv;
    }
  }
};

// One Stockham pass (decimation in frequency): N points, stride s (product of the
// radices already done), radix r.  DIR = -1 forward, +1 inverse (conjugated twiddles).
// XIn: plain pointer or WinIn; LAST: the outputs go to the overlap-add rings instead of y.
template <int DIR, class XIn, bool LAST>
__device__ __forceinline__ void fft_pass(XIn x, float2 *__restrict__ y, int N, int s, int r,
                                         const float2 *__restrict__ tw, const OlaOut &O) {
  auto put = [&](int o, float2 v) { if (LAST) O.add(o, v); else y[o] = v; };
  const int nb = N / r;             // butterflies
  const int m = nb / s;             // N / (s * r)
  if (r == 4) {
    for (int b = threadIdx.x; b < nb; b += (int)blockDim.x) {
      int p = b / s, q = b - p * s;
      const XIn xi = x + (q + s * p);
      float2 a0 = xi[0], a1 = xi[s * m], a2 = xi[2 * s * m], a3 = xi[3 * s * m];
      float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y), t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
      float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y), t3 = make_float2(a1.x - a3.x, a1.y - a3.y);
      // forward: b1 = t1 - i t3, b3 = t1 + i t3 ; inverse swaps them
      float2 mi = (DIR < 0) ? make_float2(t3.y, -t3.x) : make_float2(-t3.y, t3.x);   // (-i or +i) * t3
      float2 b0 = make_float2(t0.x + t2.x, t0.y + t2.y);
      float2 b2 = make_float2(t0.x - t2.x, t0.y - t2.y);
      float2 b1 = make_float2(t1.x + mi.x, t1.y + mi.y);
      float2 b3 = make_float2(t1.x - mi.x, t1.y - mi.y);
      const int yo = q + s * 4 * p;
      int ti = p * s;
      float2 w1 = tw[ti], w2 = tw[2 * ti], w3 = tw[3 * ti];
      if (DIR > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
      put(yo, b0); put(yo + s, cmul(b1, w1)); put(yo + 2 * s, cmul(b2, w2)); put(yo + 3 * s, cmul(b3, w3));
    }
  } else if (r == 3) {
    const float c3 = 0.86602540378443864676f;   // sin(2 pi / 3)
    for (int b = threadIdx.x; b < nb; b += (int)blockDim.x) {
      int p = b / s, q = b - p * s;
      const XIn xi = x + (q + s * p);
      const int sm = s * m;
      float2 a0 = xi[0], a1 = xi[sm], a2 = xi[2 * sm];
      float2 t = make_float2(a1.x + a2.x, a1.y + a2.y), u = make_float2(a1.x - a2.x, a1.y - a2.y);
      float2 mm = make_float2(fmaf(-0.5f, t.x, a0.x), fmaf(-0.5f, t.y, a0.y));
      // forward: -i * c3 * u ; inverse: +i * c3 * u
      float2 v = (DIR < 0) ? make_float2(c3 * u.y, -c3 * u.x) : make_float2(-c3 * u.y, c3 * u.x);
      const int yo = q + s * 3 * p;
      int ti = p * s;
      float2 w1 = tw[ti], w2 = tw[2 * ti];
      if (DIR > 0) { w1.y = -w1.y; w2.y = -w2.y; }
      put(yo, make_float2(a0.x + t.x, a0.y + t.y));
      put(yo + s, cmul(make_float2(mm.x + v.x, mm.y + v.y), w1));
      put(yo + 2 * s, cmul(make_float2(mm.x - v.x, mm.y - v.y), w2));
    }
  } else if (r == 5) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    for (int b = threadIdx.x; b < nb; b += (int)blockDim.x) {
      int p = b / s, q = b - p * s;
      const XIn xi = x + (q + s * p);
      const int sm = s * m;
      float2 a0 = xi[0], a1 = xi[sm], a2 = xi[2 * sm], a3 = xi[3 * sm], a4 = xi[4 * sm];
      float2 t1 = make_float2(a1.x + a4.x, a1.y + a4.y), t2 = make_float2(a2.x + a3.x, a2.y + a3.y);
      float2 t3 = make_float2(a1.x - a4.x, a1.y - a4.y), t4 = make_float2(a2.x - a3.x, a2.y - a3.y);
      float2 m1 = make_float2(fmaf(c1, t1.x, fmaf(c2, t2.x, a0.x)), fmaf(c1, t1.y, fmaf(c2, t2.y, a0.y)));
      float2 m2 = make_float2(fmaf(c2, t1.x, fmaf(c1, t2.x, a0.x)), fmaf(c2, t1.y, fmaf(c1, t2.y, a0.y)));
      float2 n1 = make_float2(fmaf(s1, t3.x, s2 * t4.x), fmaf(s1, t3.y, s2 * t4.y));
      float2 n2 = make_float2(fmaf(s2, t3.x, -s1 * t4.x), fmaf(s2, t3.y, -s1 * t4.y));
      // forward: b1 = m1 - i n1, b4 = m1 + i n1, b2 = m2 - i n2, b3 = m2 + i n2 (inverse: conjugate)
      float2 i1 = (DIR < 0) ? make_float2(n1.y, -n1.x) : make_float2(-n1.y, n1.x);
      float2 i2 = (DIR < 0) ? make_float2(n2.y, -n2.x) : make_float2(-n2.y, n2.x);
      const int yo = q + s * 5 * p;
      int ti = p * s;
      float2 w1 = tw[ti], w2 = tw[2 * ti], w3 = tw[3 * ti], w4 = tw[4 * ti];
      if (DIR > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; w4.y = -w4.y; }
      put(yo, make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y));
      put(yo + s, cmul(make_float2(m1.x + i1.x, m1.y + i1.y), w1));
      put(yo + 2 * s, cmul(make_float2(m2.x + i2.x, m2.y + i2.y), w2));
      put(yo + 3 * s, cmul(make_float2(m2.x - i2.x, m2.y - i2.y), w3));
      put(yo + 4 * s, cmul(make_float2(m1.x - i1.x, m1.y - i1.y), w4));
    }
  } else if (r == 8) {
    const float h = 0.70710678118654752440f;
    for (int b = threadIdx.x; b < nb; b += (int)blockDim.x) {
      int p = b / s, q = b - p * s;
      const XIn xi = x + (q + s * p);
      const int sm = s * m;
      float2 a[8];
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = xi[k * sm];
      // radix-2 stage: u_k = a_k + a_{k+4}; v_k = (a_k - a_{k+4}) w8^k  (w8 = e^{DIR i pi/4})
      float2 u[4], v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        u[k] = make_float2(a[k].x + a[k + 4].x, a[k].y + a[k + 4].y);
        v[k] = make_float2(a[k].x - a[k + 4].x, a[k].y - a[k + 4].y);
      }
      // forward w8^1 = (1 - i) h, w8^2 = -i, w8^3 = (-1 - i) h; inverse: conjugates
      if (DIR < 0) {
        v[1] = make_float2(h * (v[1].x + v[1].y), h * (v[1].y - v[1].x));
        v[2] = make_float2(v[2].y, -v[2].x);
        v[3] = make_float2(h * (v[3].y - v[3].x), -h * (v[3].x + v[3].y));
      } else {
        v[1] = make_float2(h * (v[1].x - v[1].y), h * (v[1].x + v[1].y));
        v[2] = make_float2(-v[2].y, v[2].x);
        v[3] = make_float2(-h * (v[3].x + v[3].y), h * (v[3].x - v[3].y));
      }
      // two radix-4 butterflies: even outputs from u, odd outputs from v
      float2 X[8];
#pragma unroll
      for (int par = 0; par < 2; par++) {
        const float2 *c = par ? v : u;
        float2 t0 = make_float2(c[0].x + c[2].x, c[0].y + c[2].y), t1 = make_float2(c[0].x - c[2].x, c[0].y - c[2].y);
        float2 t2 = make_float2(c[1].x + c[3].x, c[1].y + c[3].y), t3 = make_float2(c[1].x - c[3].x, c[1].y - c[3].y);
        float2 mi = (DIR < 0) ? make_float2(t3.y, -t3.x) : make_float2(-t3.y, t3.x);
        X[par] = make_float2(t0.x + t2.x, t0.y + t2.y);
        X[4 + par] = make_float2(t0.x - t2.x, t0.y - t2.y);
        X[2 + par] = make_float2(t1.x + mi.x, t1.y + mi.y);
        X[6 + par] = make_float2(t1.x - mi.x, t1.y - mi.y);
      }
      const int yo = q + s * 8 * p;
      const int ti = p * s;
      put(yo, X[0]);
#pragma unroll
      for (int j = 1; j < 8; j++) {
        float2 w = tw[j * ti];
        if (DIR > 0) w.y = -w.y;
        put(yo + j * s, cmul(X[j], w));
      }
    }
  } else if (r == 2) {
    for (int b = threadIdx.x; b < nb; b += (int)blockDim.x) {
      int p = b / s, q = b - p * s;
      float2 a0 = x[q + s * p], a1 = x[q + s * (p + m)];
      float2 w1 = tw[p * s];
      if (DIR > 0) w1.y = -w1.y;
      put(q + s * 2 * p, make_float2(a0.x + a1.x, a0.y + a1.y));
      put(q + s * (2 * p + 1), cmul(make_float2(a0.x - a1.x, a0.y - a1.y), w1));
    }
  } else {
    // generic radix: one output per work item, omega_r^(jk) looked up in the N-point table
    const int wstep = N / r;
    for (int idx = threadIdx.x; idx < N; idx += (int)blockDim.x) {
      int j = idx / nb, b = idx - j * nb;
      int p = b / s, q = b - p * s;
      const XIn xi = x + (q + s * p);
      const int sm = s * m;
      float2 acc = xi[0];
      int t = 0;
      const int st = j * wstep;
      for (int k = 1; k < r; k++) {
        t += st;
        if (t >= N) t -= N;
        float2 w = tw[t];
        if (DIR > 0) w.y = -w.y;
        float2 a = xi[k * sm];
        acc.x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc.x));
        acc.y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc.y));
      }
      float2 w2 = tw[p * j * s];
      if (DIR > 0) w2.y = -w2.y;
      put(q + s * (r * p + j), cmul(acc, w2));
    }
  }
}

// full transform; returns the buffer holding the result.  FIN: the first pass reads the windowed staging
// buffer (W); FOUT: the last pass adds into the overlap-add rings (O) and nothing is written to the buffer.
template <int DIR, bool FIN, bool FOUT>
__device__ float2 *fft_run(float2 *a, float2 *b, const FftPlan &pl, const float2 *tw, const WinIn &W, const OlaOut &O) {
  int s = 1;
  float2 *src = a, *dst = b;
  for (int i = 0; i < pl.npass; i++) {
    const int r = pl.radix[i];
    const bool first = FIN && i == 0, last = FOUT && i == pl.npass - 1;
    if (first && last) fft_pass<DIR, WinIn, true>(W, dst, pl.n, s, r, tw, O);
    else if (first) fft_pass<DIR, WinIn, false>(W, dst, pl.n, s, r, tw, O);
    else if (last) fft_pass<DIR, const float2 *, true>(src, dst, pl.n, s, r, tw, O);
    else fft_pass<DIR, const float2 *, false>(src, dst, pl.n, s, r, tw, O);
    __syncthreads();
    s *= r;
    float2 *t = src; src = dst; dst = t;
  }
  return src;
}

// Compile-time plans for soundgen's usual windows (50 ms at 16 / 22.05 / 24 / 44.1 / 48 kHz and the
// 10 ms presets): with N, the strides and the radices known, every index division folds to a
// multiply-shift and the radix dispatch disappears.  SPEC 0 = run-time plan (any even length).
template <int DIR, bool FIN, bool FOUT, int N, int S, int R, int... Rest>
__device__ __forceinline__ float2 *run_ct(float2 *a, float2 *b, const float2 *tw, const WinIn &W, const OlaOut &O) {
  constexpr bool last = sizeof...(Rest) == 0;
  if constexpr (S == 1 && FIN) fft_pass<DIR, WinIn, (last && FOUT)>(W, b, N, S, R, tw, O);
  else fft_pass<DIR, const float2 *, (last && FOUT)>(a, b, N, S, R, tw, O);
  __syncthreads();
  if constexpr (last) return b;
  else return run_ct<DIR, FIN, FOUT, N, S * R, Rest...>(b, a, tw, W, O);
}
template <int DIR, int SPEC, bool FIN, bool FOUT>
__device__ __forceinline__ float2 *fft_any(float2 *a, float2 *b, const FftPlan &pl, const float2 *tw, const WinIn &W,
                                           const OlaOut &O) {
  if constexpr (SPEC == 1) return run_ct<DIR, FIN, FOUT, 800, 1, 5, 5, 4, 8>(a, b, tw, W, O);
  else if constexpr (SPEC == 2) return run_ct<DIR, FIN, FOUT, 1102, 1, 29, 19, 2>(a, b, tw, W, O);
  else if constexpr (SPEC == 3) return run_ct<DIR, FIN, FOUT, 1200, 1, 5, 5, 3, 4, 4>(a, b, tw, W, O);
  else if constexpr (SPEC == 4) return run_ct<DIR, FIN, FOUT, 2204, 1, 29, 19, 4>(a, b, tw, W, O);
  else if constexpr (SPEC == 5) return run_ct<DIR, FIN, FOUT, 2400, 1, 5, 5, 3, 4, 8>(a, b, tw, W, O);
  else if constexpr (SPEC == 6) return run_ct<DIR, FIN, FOUT, 160, 1, 5, 4, 8>(a, b, tw, W, O);
  else return fft_run<DIR, FIN, FOUT>(a, b, pl, tw, W, O);
}

__device__ __forceinline__ int frame_in_start(const FftPlan &pl, int k) {   // 0-based
  return (int)floor(1.0 + (double)k * pl.h_in) - 1;
}
__device__ __forceinline__ int frame_out_start(const FftPlan &pl, int k) {
  return (int)floor((double)k * pl.h_out + 1.0) - 1;
}

// MODE 0: filter (K2).  MODE 1: noise (K5), UT = uniform dtype.
template <int MODE, typename UT, int SPEC>
__global__ void __launch_bounds__(FFT_THREADS, 2)
k_stft(const FftSeg *__restrict__ segs, const FftJob *__restrict__ jobs, const FftPlan *__restrict__ plans,
       const float2 *__restrict__ twpool, const float *__restrict__ winpool,
       const float *__restrict__ in_f, const UT *__restrict__ in_u, const float *__restrict__ envpool,
       float *__restrict__ outpool, int *__restrict__ maxpool) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t bars[2];
  __shared__ float red[FFT_THREADS / 32];

  const FftSeg sg = segs[blockIdx.x];
  const FftJob jb = jobs[sg.job];
  const FftPlan pl = plans[jb.plan];
  const int N = pl.n, nr = N / 2;
  const int hceil = (int)ceil(pl.h_out) + 2;
  const int ring = N + 2 * hceil + 4;
  const int stage_len = (N + (int)ceil(pl.h_in) + 12) & ~3;     // floats per staging buffer

  float2 *bufA = reinterpret_cast<float2 *>(smem_raw);
  float2 *bufB = bufA + N;
  float2 *tw = bufB + N;
  float *ola = reinterpret_cast<float *>(tw + N);      // ring A: even frames of the pairs
  float *olb = ola + ((ring + 3) & ~3);                // ring B: odd frames
  float *stage0 = olb + ((ring + 3) & ~3);
  float *stage1 = stage0 + stage_len;
  float *vec = stage1 + ((MODE == 0) ? stage_len : 0);           // noise: rolloff vector [nr]

  const float2 *twg = twpool + pl.tw_off;
  const float *wa = winpool + pl.wa_off;
  const float *ws = winpool + pl.ws_off;
  for (int i = threadIdx.x; i < N; i += (int)blockDim.x) tw[i] = twg[i];
  for (int i = threadIdx.x; i < ring; i += (int)blockDim.x) { ola[i] = 0.0f; olb[i] = 0.0f; }
  if (MODE == 1) {
    // rolloff vector 2^(rolloffNoise/10 * log2(1:nr)) (source.R:103-105)
    for (int i = threadIdx.x; i < nr; i += (int)blockDim.x)
      vec[i] = (float)exp2(jb.rolloffNoise / 10.0 * log2((double)(i + 1)));
  }
  if (MODE == 0 && threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nwarm = (int)ceil((double)N / pl.h_out) - 1;
  const int kstart = max(0, sg.ka - nwarm);
  const int flush_lo = (sg.ka == 0) ? 0 : frame_out_start(pl, sg.ka);
  const int flush_hi = (sg.kb >= jb.nc) ? jb.xlen : frame_out_start(pl, sg.kb);
  int flushed = frame_out_start(pl, kstart);      // ring position accounted so far
  if (sg.ka == 0) flushed = 0;
  const float *src_f = (MODE == 0) ? (in_f + jb.in_off) : nullptr;
  float vmax = -INFINITY;

  // TMA prefetch of the first pair
  uint32_t phase[2] = {0u, 0u};
  auto issue_load = [&](int k, int buf) {
    // frames k, k+1 need input [s_k, s_{k+1} + N)
    int sA = frame_in_start(pl, k);
    int kB = min(k + 1, jb.nc - 1);
    int sB = frame_in_start(pl, kB);
    int a0 = sA & ~3;
    int cnt = ((sB + N - a0) + 3) & ~3;
    if (cnt > stage_len) cnt = stage_len;
    uint32_t bytes = (uint32_t)cnt * 4u;
    mbar_expect_tx(&bars[buf], bytes);
    tma_load_1d(buf ? stage1 : stage0, src_f + a0, bytes, &bars[buf]);
  };
  if (MODE == 0 && threadIdx.x == 0 && kstart < sg.kb) issue_load(kstart, 0);

  int cur = 0;
  for (int k = kstart; k < sg.kb; k += 2) {
    const bool hasB = (k + 1 < sg.kb) && (k + 1 < jb.nc);
    float2 *spec;
    if (MODE == 0) {
      // ---- wait for the staged input, prefetch the next pair ----
      if (!mbar_wait(&bars[cur], phase[cur])) return;   // the host reports it (stft_timeout_flag)
      phase[cur] ^= 1u;
      if (threadIdx.x == 0 && k + 2 < sg.kb) issue_load(k + 2, cur ^ 1);
      const float *st = cur ? stage1 : stage0;
      const int sA = frame_in_start(pl, k);
      const int a0 = sA & ~3;
      WinIn W;
      W.a = st + (sA - a0); W.b = st + (hasB ? (frame_in_start(pl, k + 1) - a0) : 0); W.w = wa; W.off = 0; W.hasB = hasB;
      OlaOut O0 = {};
      // Register-butterfly first passes (radix 2/3/4/5/8) read every input once: there the analysis window
      // and the A + iB packing are fused into the pass.  A generic (prime) first radix reads every input r
      // times, so those plans (1102 = 29*19*2, 2204, clamped windows) window the frame in a pass of its own.
      constexpr bool FUSE_CT = (SPEC == 1 || SPEC == 3 || SPEC == 5 || SPEC == 6);
      const bool fuse_rt = (SPEC == 0) && (pl.radix[0] <= 5 || pl.radix[0] == 8);
      float2 *Z;
      if (FUSE_CT || fuse_rt) {
        Z = fft_any<-1, SPEC, true, false>(bufA, bufB, pl, tw, W, O0);
      } else {
        for (int i = threadIdx.x; i < N; i += (int)blockDim.x) bufA[i] = W[i];
        __syncthreads();
        Z = fft_any<-1, SPEC, false, false>(bufA, bufB, pl, tw, W, O0);
      }
      // all generic-proxy reads of this staging buffer are done: make it safe for the next TMA write
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // ---- split the two spectra, multiply by the envelope, rebuild Hermitian halves ----
      const float *eA = envpool + jb.env_off + (int64_t)((jb.nint > 1) ? k : 0) * nr;
      const float *eB = envpool + jb.env_off + (int64_t)((jb.nint > 1 && hasB) ? (k + 1) : 0) * nr;
      for (int kk = threadIdx.x; kk < nr; kk += (int)blockDim.x) {
        if (kk == 0) {
          float2 z0 = Z[0];
          Z[0] = make_float2(eA[0] * z0.x, eB[0] * z0.y);
        } else {
          float2 zk = Z[kk], zm = Z[N - kk];
          // XA = (zk + conj(zm))/2 ; XB = (zk - conj(zm))/(2i)
          float2 XA = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
          float2 XB = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
          float ea = eA[kk], eb = eB[kk];
          float2 FA = make_float2(ea * XA.x, ea * XA.y), FB = make_float2(eb * XB.x, eb * XB.y);
          Z[kk] = make_float2(FA.x - FB.y, FA.y + FB.x);            // FA + i FB
          Z[N - kk] = make_float2(FA.x + FB.y, FB.x - FA.y);        // conj(FA) + i conj(FB)
          if (kk == nr - 1) Z[nr] = make_float2(FA.x, FB.x);        // Nyquist := Re(last bin) (seewave.r:3474)
        }
      }
      __syncthreads();
      spec = Z;
    } else {
      // ---- noise: real, zero-phase spectrum u * filter (source.R:111-114) ----
      const UT *uA = in_u + jb.in_off + (int64_t)k * nr;
      const UT *uB = in_u + jb.in_off + (int64_t)(k + 1) * nr;
      const float *fA = nullptr, *fB = nullptr;
      if (jb.env_off >= 0) {
        // filterRowIdx = round(seq(1, ncol, length.out = nc)) (source.R:99)
        int cA = (int)rint(r_seq_at(1.0, (double)jb.nint, jb.nc, k)) - 1;
        int cB = hasB ? (int)rint(r_seq_at(1.0, (double)jb.nint, jb.nc, k + 1)) - 1 : 0;
        fA = envpool + jb.env_off + (int64_t)cA * nr;
        fB = envpool + jb.env_off + (int64_t)cB * nr;
      }
      for (int kk = threadIdx.x; kk < nr; kk += (int)blockDim.x) {
        float ro = vec[kk];
        float va = (float)uA[kk] * ro * (fA ? fA[kk] : 1.0f);
        float vb = hasB ? (float)uB[kk] * ro * (fB ? fB[kk] : 1.0f) : 0.0f;
        float2 v = make_float2(va, vb);
        bufA[kk] = v;
        if (kk > 0) bufA[N - kk] = v;
        if (kk == nr - 1) bufA[nr] = v;
      }
      __syncthreads();
      spec = bufA;
    }
    float2 *other = (spec == bufA) ? bufB : bufA;
    // the last inverse pass adds the synthesis-windowed frames straight into the two rings
    OlaOut O;
    O.ra = ola; O.rb = olb; O.ws = ws; O.ring = ring; O.hasB = hasB;
    O.oA = frame_out_start(pl, k) % ring;
    O.oB = hasB ? frame_out_start(pl, k + 1) % ring : 0;
    WinIn W0 = {};
    fft_any<+1, SPEC, false, true>(spec, other, pl, tw, W0, O);
    // ---- samples before the next frame's start are final: write them out once ----
    int knext = k + 2;
    int done_to = (knext < sg.kb) ? frame_out_start(pl, knext) : ((sg.kb >= jb.nc) ? jb.xlen : frame_out_start(pl, sg.kb));
    if (knext >= sg.kb && sg.kb < jb.nc) done_to = flush_hi;
    const int fbase = flushed % ring;
    for (int t = flushed + threadIdx.x; t < done_to; t += (int)blockDim.x) {
      int slot = fbase + (t - flushed);
      if (slot >= ring) slot -= ring;
      float v = ola[slot] + olb[slot];
      ola[slot] = 0.0f; olb[slot] = 0.0f;
      if (t >= flush_lo && t < flush_hi) {
        int oi = t - jb.shift;
        if (oi >= 0 && oi < jb.out_len) {
          outpool[jb.out_off + oi] = v;
          vmax = fmaxf(vmax, v);
        }
      }
    }
    flushed = max(flushed, done_to);
    __syncthreads();
    cur ^= 1;
  }

  // ---- signed maximum of what this CTA emitted ----
  for (int of = 16; of > 0; of >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, of));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = vmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); i++) v = fmaxf(v, red[i]);
    if (v > -INFINITY) atomicMax(&maxpool[jb.max_slot], float_to_ordered(v));
  }
}

size_t stft_smem_bytes(int n, double h_in, double h_out, int mode) {
  int hceil = (int)ceil(h_out) + 2;
  int ring = n + 2 * hceil + 4;
  int stage_len = (n + (int)ceil(h_in) + 12) & ~3;
  size_t floats = (size_t)6 * n + 2 * (size_t)((ring + 3) & ~3) + (mode == 0 ? 2 * (size_t)stage_len : (size_t)stage_len + n / 2 + 4);
  return floats * 4 + 64;
}

#define N_SPEC 7
static size_t attr_smem[3][N_SPEC];

template <typename K>
static cudaError_t ensure_smem(K kf, size_t *slot, size_t smem) {
  if (smem <= 48 * 1024 || smem <= *slot) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) *slot = smem;
  return e;
}

int stft_spec_of(int n, const int *radix, int npass) {
  static const int sizes[N_SPEC] = {0, 800, 1102, 1200, 2204, 2400, 160};
  static const int rad[N_SPEC][8] = {{0}, {5, 5, 4, 8}, {29, 19, 2}, {5, 5, 3, 4, 4}, {29, 19, 4},
                                     {5, 5, 3, 4, 8}, {5, 4, 8}};
  for (int sp = 1; sp < N_SPEC; sp++) {
    if (sizes[sp] != n) continue;
    int np = 0;
    while (np < 8 && rad[sp][np]) np++;
    if (np != npass) continue;
    bool same = true;
    for (int i = 0; i < np; i++) if (rad[sp][i] != radix[i]) same = false;
    if (same) return sp;
  }
  return 0;
}

template <int SPEC>
static cudaError_t launch_spec(int mode, int u_is_float, const FftSeg *segs, int n_segs, const FftJob *jobs,
                               const FftPlan *plans, const float2 *tw, const float *win, const float *in_f,
                               const void *in_u, const float *env, float *out, int *maxpool, size_t smem,
                               cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (mode == 0) {
    auto kf = k_stft<0, float, SPEC>;
    if ((e = ensure_smem(kf, &attr_smem[0][SPEC], smem)) != cudaSuccess) return e;
    kf<<<n_segs, (SPEC == 0 || SPEC == 2 || SPEC == 4) ? 256 : FFT_THREADS, smem, st>>>(segs, jobs, plans, tw, win, in_f, nullptr, env, out, maxpool);
  } else if (u_is_float) {
    auto kf = k_stft<1, float, SPEC>;
    if ((e = ensure_smem(kf, &attr_smem[1][SPEC], smem)) != cudaSuccess) return e;
    kf<<<n_segs, (SPEC == 0 || SPEC == 2 || SPEC == 4) ? 256 : FFT_THREADS, smem, st>>>(segs, jobs, plans, tw, win, nullptr, (const float *)in_u, env, out, maxpool);
  } else {
    auto kf = k_stft<1, double, SPEC>;
    if ((e = ensure_smem(kf, &attr_smem[2][SPEC], smem)) != cudaSuccess) return e;
    kf<<<n_segs, (SPEC == 0 || SPEC == 2 || SPEC == 4) ? 256 : FFT_THREADS, smem, st>>>(segs, jobs, plans, tw, win, nullptr, (const double *)in_u, env, out, maxpool);
  }
  return cudaGetLastError();
}

// all segments of one launch share `spec` (0 = run-time plan)
cudaError_t launch_stft(int mode, int u_is_float, int spec, const FftSeg *segs, int n_segs, const FftJob *jobs,
                        const FftPlan *plans, const float2 *tw, const float *win, const float *in_f,
                        const void *in_u, const float *env, float *out, int *maxpool, size_t smem,
                        cudaStream_t st) {
  if (n_segs <= 0) return cudaSuccess;
#define SGB_CASE(SP) case SP: return launch_spec<SP>(mode, u_is_float, segs, n_segs, jobs, plans, tw, win, in_f, in_u, env, out, maxpool, smem, st);
  switch (spec) {
    SGB_CASE(1) SGB_CASE(2) SGB_CASE(3) SGB_CASE(4) SGB_CASE(5) SGB_CASE(6)
    default: return launch_spec<0>(mode, u_is_float, segs, n_segs, jobs, plans, tw, win, in_f, in_u, env, out, maxpool, smem, st);
  }
#undef SGB_CASE
}
