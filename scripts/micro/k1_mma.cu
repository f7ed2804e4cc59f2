// Microbenchmark for VERDICT r01 item 7: is the additive-synthesis hot loop (R/source.R:396-419) a
// contraction after all?  With rows j = K b + m,
//   sum_j a_j(u) sin(j th_u) = sum_b [ sin(K b th_u) C_b(u) + cos(K b th_u) S_b(u) ],
//   C_b(u) = sum_m a_{Kb+m}(u) cos(m th_u),  S_b(u) = sum_m a_{Kb+m}(u) sin(m th_u),  a = Y + w(u) dY,
// so C and S are small GEMMs [blocks x K] . [K x samples] against a trig matrix that costs K values per
// sample instead of one recurrence step per (row, sample).  Here: K = 16, one warp per 32-sample tile,
// mma.sync.m16n8k8 TF32 with the 3xTF32 split (hi*hi + lo*hi + hi*lo) for ~FP32 accuracy, rotators
// exp(i K b th) advanced per 16-block chunk in the epilogue.  Measures partial-samples per second and the
// error against an FP64 direct sum.     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o k1_mma k1_mma.cu
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int KR = 16;              // rows per block
constexpr int CH = 16;              // blocks per chunk (the M of the MMA)
constexpr double TWO_PI = 6.283185307179586;

__device__ __forceinline__ uint32_t tf32_hi(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void split(float x, uint32_t &hi, uint32_t &lo) { hi = tf32_hi(x); lo = tf32_hi(x - __uint_as_float(hi)); }
__device__ __forceinline__ void mma(float *d, const uint32_t *a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tab: per column (glottal cycle) J4 rows of float4 {Y_hi, Y_lo, dY_hi, dY_lo} (tf32 split done by the producer).
// tile t: 32 samples with phases th[t*32 + u] (cycles, double), weights w[], amplitude column col[t].
__global__ void __launch_bounds__(128) k_mma(const float4 *__restrict__ tab, int Jpad, const double *__restrict__ th,
                                             const float *__restrict__ w, const int *__restrict__ col, int ntiles,
                                             float *__restrict__ out) {
  const int lane = threadIdx.x & 31, t = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (t >= ntiles) return;
  const int r = lane >> 2, q = lane & 3;
  const float4 *T = tab + (size_t)col[t] * Jpad;
  const double *tht = th + (size_t)t * 32;
  // ---- trig B fragments: B[m][n], m = q (+4) + 8 ks, sample n = r + 8 nt; cos and sin, hi / lo ----
  uint32_t bc[2][4][2][2], bs[2][4][2][2];      // [ks][nt][reg][hi/lo]
#pragma unroll
  for (int nt = 0; nt < 4; nt++) {
    const double thn = tht[r + 8 * nt];
    const double fr = thn - floor(thn);
    float s1, c1;
    sincospif(2.0f * (float)fr, &s1, &c1);            // e^{i th}: th reduced to [0, 1) cycles in FP64 first
    // e^{i q th} and e^{i 4 th}
    float cq = 1.f, sq = 0.f;
    for (int i = 0; i < q; i++) { float c = cq * c1 - sq * s1, s = cq * s1 + sq * c1; cq = c; sq = s; }
    float c2 = c1 * c1 - s1 * s1, s2 = 2.f * c1 * s1;
    float c4 = c2 * c2 - s2 * s2, s4 = 2.f * c2 * s2;
    float cm = cq, sm = sq;
#pragma unroll
    for (int j = 0; j < 4; j++) {                     // m = q + 4 j: ks = j / 2, reg = j % 2
      split(cm, bc[j >> 1][nt][j & 1][0], bc[j >> 1][nt][j & 1][1]);
      split(sm, bs[j >> 1][nt][j & 1][0], bs[j >> 1][nt][j & 1][1]);
      float c = cm * c4 - sm * s4, s = cm * s4 + sm * c4; cm = c; sm = s;
    }
  }
  // ---- epilogue state: samples u = 8 nt + 2 q + e, blocks b = 16 c + r (+8) ----
  float rc[4][2][2], rs[4][2][2], stc[4][2], sts[4][2], ww[4][2], acc[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int u = 8 * nt + 2 * q + e;
      const double thu = tht[u];
      ww[nt][e] = w[(size_t)t * 32 + u];
      acc[nt][e] = 0.f;
      double a = thu * (double)(KR * CH); a -= floor(a);
      sincospif(2.0f * (float)a, &sts[nt][e], &stc[nt][e]);          // chunk step e^{i K CH th}
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double p = thu * (double)(KR * (r + 8 * h)); p -= floor(p);   // e^{i K b th}, b = r + 8 h
        sincospif(2.0f * (float)p, &rs[nt][e][h], &rc[nt][e][h]);
      }
    }
  const int nchunks = Jpad / (KR * CH);
  for (int c = 0; c < nchunks; c++) {
    float D[4][4][4];                                  // [kind: Ycos, dYcos, Ysin, dYsin][nt][4]
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int i = 0; i < 4; i++) D[k][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      // A fragments: rows b = 16 c + r (+8), cols m = 8 ks + q (+4): table index j = KR b + m
      uint32_t ay[2][4], ad[2][4];                     // [hi/lo][a0..a3]
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int b = CH * c + r + 8 * (i & 1), m = 8 * ks + q + 4 * (i >> 1);
        const float4 v = __ldg(&T[KR * b + m]);
        ay[0][i] = __float_as_uint(v.x); ay[1][i] = __float_as_uint(v.y);
        ad[0][i] = __float_as_uint(v.z); ad[1][i] = __float_as_uint(v.w);
      }
      // 3xTF32: hi*hi + lo*hi + hi*lo.  The three products of one accumulator are issued 16 MMAs apart
      // (all 16 accumulator tiles per split term), so that no MMA waits on the one before it.
#pragma unroll
      for (int sp = 0; sp < 3; sp++) {
        const int ah = (sp == 1) ? 1 : 0, bh = (sp == 2) ? 1 : 0;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          mma(D[0][nt], ay[ah], bc[ks][nt][0][bh], bc[ks][nt][1][bh]);
          mma(D[1][nt], ad[ah], bc[ks][nt][0][bh], bc[ks][nt][1][bh]);
          mma(D[2][nt], ay[ah], bs[ks][nt][0][bh], bs[ks][nt][1][bh]);
          mma(D[3][nt], ad[ah], bs[ks][nt][0][bh], bs[ks][nt][1][bh]);
        }
      }
    }
    // epilogue: D[.][nt][2 h + e] belongs to block row r + 8 h, sample 8 nt + 2 q + e
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int i = 2 * h + e;
          const float C = fmaf(ww[nt][e], D[1][nt][i], D[0][nt][i]), S = fmaf(ww[nt][e], D[3][nt][i], D[2][nt][i]);
          float &cr = rc[nt][e][h], &sr = rs[nt][e][h];
          acc[nt][e] = fmaf(sr, C, fmaf(cr, S, acc[nt][e]));
          const float cn = cr * stc[nt][e] - sr * sts[nt][e], sn = cr * sts[nt][e] + sr * stc[nt][e];
          cr = cn; sr = sn;
        }
  }
#pragma unroll
  for (int nt = 0; nt < 4; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      float v = acc[nt][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (r == 0) out[(size_t)t * 32 + 8 * nt + 2 * q + e] = v;
    }
}

// the reference's formulation in FP64: one thread per sample
__global__ void k_ref(const double *Y, const double *dY, int J, int Jpad, const double *th, const float *w, const int *col,
                      int nsamp, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsamp) return;
  const double *y = Y + (size_t)col[i / 32] * Jpad, *d = dY + (size_t)col[i / 32] * Jpad;
  double s = 0, ww = w[i], t = th[i];
  for (int j = 1; j < J; j++) s += (y[j] + ww * d[j]) * sin(TWO_PI * (double)j * t);
  out[i] = s;
}

int main(int argc, char **argv) {
  const int J = argc > 1 ? atoi(argv[1]) : 1024;          // rows incl. j = 0
  const int ntiles = argc > 2 ? atoi(argv[2]) : 1 << 18;  // 32-sample tiles
  const int ncol = 4096, Jpad = ((J + KR * CH - 1) / (KR * CH)) * (KR * CH);
  std::vector<double> Y((size_t)ncol * Jpad, 0.0), dY((size_t)ncol * Jpad, 0.0), th((size_t)ntiles * 32);
  std::vector<float4> tab((size_t)ncol * Jpad);
  std::vector<float> w((size_t)ntiles * 32);
  std::vector<int> col(ntiles);
  srand(1);
  auto rnd = [] { return rand() / (double)RAND_MAX; };
  for (int c = 0; c < ncol; c++) {
    const double slope = -1.0 - 5.0 * rnd();              // dB per octave, cfg3's range
    for (int j = 1; j < J; j++) {
      double a = pow(2.0, slope * log2((double)j) / 10.0) * (0.7 + 0.6 * rnd());
      Y[(size_t)c * Jpad + j] = a; dY[(size_t)c * Jpad + j] = a * 0.2 * (rnd() - 0.5);
    }
    for (int j = 0; j < Jpad; j++) {
      auto sp = [](double x, float &hi, float &lo) { float f = (float)x; uint32_t b; memcpy(&b, &f, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&hi, &b, 4); float rr = f - hi; memcpy(&b, &rr, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&lo, &b, 4); };
      float4 v; sp(Y[(size_t)c * Jpad + j], v.x, v.y); sp(dY[(size_t)c * Jpad + j], v.z, v.w); tab[(size_t)c * Jpad + j] = v;
    }
  }
  for (int t = 0; t < ntiles; t++) {
    col[t] = rand() % ncol;
    double f0 = 50 + 70 * rnd(), ph = 1000 * rnd();       // cycles: phase keeps growing along a syllable
    for (int u = 0; u < 32; u++) { th[(size_t)t * 32 + u] = (ph + u * f0 / 48000.0) / 5.0; w[(size_t)t * 32 + u] = (float)rnd(); }
  }
  float4 *dtab; double *dth, *dY64, *ddY64, *dref; float *dw, *dout; int *dcol;
  CK(cudaMalloc(&dtab, tab.size() * 16)); CK(cudaMalloc(&dth, th.size() * 8)); CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMalloc(&dcol, col.size() * 4)); CK(cudaMalloc(&dout, w.size() * 4));
  CK(cudaMemcpy(dtab, tab.data(), tab.size() * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dth, th.data(), th.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcol, col.data(), col.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) k_mma<<<(ntiles + 3) / 4, 128>>>(dtab, Jpad, dth, dw, dcol, ntiles, dout);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int i = 0; i < reps; i++) k_mma<<<(ntiles + 3) / 4, 128>>>(dtab, Jpad, dth, dw, dcol, ntiles, dout);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  const double ps = (double)ntiles * 32 * (J - 1);
  printf("J %d (padded %d), %d tiles: %.3f ms, %.3e partial-samples/s, %.2f TFLOP/s at 6 flop per partial-sample\n", J, Jpad, ntiles, ms, ps / (ms * 1e-3), 6 * ps / (ms * 1e-3) / 1e12);
  // accuracy on the first 2048 tiles
  const int nchk = 2048 * 32;
  CK(cudaMalloc(&dY64, Y.size() * 8)); CK(cudaMalloc(&ddY64, dY.size() * 8)); CK(cudaMalloc(&dref, nchk * 8));
  CK(cudaMemcpy(dY64, Y.data(), Y.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ddY64, dY.data(), dY.size() * 8, cudaMemcpyHostToDevice));
  k_ref<<<(nchk + 127) / 128, 128>>>(dY64, ddY64, J, Jpad, dth, dw, dcol, nchk, dref);
  std::vector<double> ref(nchk); std::vector<float> got(nchk);
  CK(cudaMemcpy(ref.data(), dref, nchk * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(got.data(), dout, nchk * 4, cudaMemcpyDeviceToHost));
  double peak = 0, worst = 0;
  for (int i = 0; i < nchk; i++) { peak = fmax(peak, fabs(ref[i])); worst = fmax(worst, fabs(ref[i] - got[i])); }
  printf("max |err| = %.3e of peak %.3f (bar: 1e-4)\n", worst / peak, peak);
  return 0;
}
