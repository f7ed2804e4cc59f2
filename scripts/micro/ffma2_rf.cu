// micro-benchmark: FFMA2 issue rate on sm_100a as a function of how many distinct register
// operands an instruction reads (register-file bandwidth / operand reuse), and of the K1 inner
// loop formulations.  Prints packed FMA instructions per cycle per SMSP (peak 0.5).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// A: one varying operand (the other two warp-uniform constants)
__global__ void __launch_bounds__(128, 6) kA(float2 *out, int it) {
  float2 a[8];
  const float2 m = f2(1.0000001f, 0.9999999f), c = f2(1e-9f, -1e-9f);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = f2(1.0f + i + threadIdx.x * 1e-3f, 2.0f + i);
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], m, c);
  }
  float2 s = a[0];
#pragma unroll
  for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
  if (s.x == 123.456f) out[threadIdx.x] = s;
}
// B: three distinct per-lane register operands, no sharing between neighbours
__global__ void __launch_bounds__(128, 6) kB(float2 *out, int it) {
  float2 a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = f2(1.0f + i + threadIdx.x * 1e-3f, 2.0f + i);
    b[i] = f2(0.999f + i * 1e-5f + threadIdx.x * 1e-7f, 0.998f);
    c[i] = f2(1e-3f * (i + 1) + threadIdx.x * 1e-7f, 1e-4f);
  }
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], b[i], c[i]);
  }
  float2 s = a[0];
#pragma unroll
  for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
  if (s.x == 123.456f) out[threadIdx.x] = s;
}
// C: per-lane operands, but neighbours share the multiplier (slot reuse possible)
__global__ void __launch_bounds__(128, 6) kC(float2 *out, int it) {
  float2 a[8], c[8];
  float2 b = f2(0.999f + threadIdx.x * 1e-7f, 0.998f);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = f2(1.0f + i + threadIdx.x * 1e-3f, 2.0f + i);
    c[i] = f2(1e-3f * (i + 1) + threadIdx.x * 1e-7f, 1e-4f);
  }
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], b, c[i]);
  }
  float2 s = a[0];
#pragma unroll
  for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
  if (s.x == 123.456f) out[threadIdx.x] = s;
}
// C2: neighbours share multiplier AND addend (only the accumulator is fresh)
__global__ void __launch_bounds__(128, 6) kC2(float2 *out, int it) {
  float2 a[8];
  float2 b = f2(0.999f + threadIdx.x * 1e-7f, 0.998f), c = f2(1e-3f + threadIdx.x * 1e-7f, 1e-4f);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = f2(1.0f + i + threadIdx.x * 1e-3f, 2.0f + i);
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], b, c);
  }
  float2 s = a[0];
#pragma unroll
  for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
  if (s.x == 123.456f) out[threadIdx.x] = s;
}
// S: scalar FFMA, three distinct register operands
__global__ void __launch_bounds__(128, 6) kS(float2 *out, int it) {
  float a[16], b[16], c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    a[i] = 1.0f + i + threadIdx.x * 1e-3f; b[i] = 0.999f + i * 1e-5f + threadIdx.x * 1e-7f;
    c[i] = 1e-3f * (i + 1) + threadIdx.x * 1e-7f;
  }
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b[i], c[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  if (s == 123.456f) out[threadIdx.x] = f2(s, s);
}

// ---- K1 inner-loop formulations: 64 rows per block, amplitudes {Y,Y,dY,dY} broadcast from smem ----
#define ROWS 64
template <int V>
__global__ void __launch_bounds__(128, 6) kL(float2 *out, int nblk) {
  __shared__ float4 tab[2 * ROWS];
  for (int i = threadIdx.x; i < 2 * ROWS; i += blockDim.x) tab[i] = make_float4(1e-3f * i, 1e-3f * i, 1e-5f, 1e-5f);
  __syncthreads();
  float2 w2[2], delta2[2], sigma2[2];
  for (int p = 0; p < 2; p++) {
    w2[p] = f2(0.1f + threadIdx.x * 1e-3f, 0.2f + p);
    delta2[p] = f2(-1e-3f - threadIdx.x * 1e-6f, -2e-3f);
    sigma2[p] = f2(1.0f, (threadIdx.x & 1) ? -1.0f : 1.0f);
  }
  float2 w4[4], d4[4];
  for (int p = 0; p < 4; p++) { w4[p] = f2(0.1f + threadIdx.x * 1e-3f, 0.2f + p); d4[p] = f2(-1e-3f - threadIdx.x * 1e-6f, -2e-3f - p * 1e-4f); }
  float2 acc[4] = {f2(0, 0), f2(0, 0), f2(0, 0), f2(0, 0)};
  for (int b = 0; b < nblk; b++) {
    if (V == 0) {          // general: per-lane sigma, 2 chains, 4 FFMA2 per pair-row
      float2 bb[2] = {f2(0, 0), f2(0, 0)}, dd[2] = {f2(0, 0), f2(0, 0)};
#pragma unroll 8
      for (int m = ROWS - 1; m >= 0; m--) {
        const float4 q = tab[m];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float2 a_ = __ffma2_rn(w2[c], f2(q.z, q.w), f2(q.x, q.y));
          dd[c] = __ffma2_rn(delta2[c], bb[c], __ffma2_rn(sigma2[c], dd[c], a_));
          bb[c] = __ffma2_rn(sigma2[c], bb[c], dd[c]);
        }
      }
      acc[0].x += bb[0].x + dd[0].y; acc[1].x += bb[1].x + dd[1].y;
    } else if (V == 1) {   // sigma = +1 for the whole warp: adds instead of two of the FMAs
      float2 bb[2] = {f2(0, 0), f2(0, 0)}, dd[2] = {f2(0, 0), f2(0, 0)};
#pragma unroll 8
      for (int m = ROWS - 1; m >= 0; m--) {
        const float4 q = tab[m];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float2 a_ = __ffma2_rn(w2[c], f2(q.z, q.w), f2(q.x, q.y));
          dd[c] = __ffma2_rn(delta2[c], bb[c], __fadd2_rn(dd[c], a_));
          bb[c] = __fadd2_rn(bb[c], dd[c]);
        }
      }
      acc[0].x += bb[0].x + dd[0].y; acc[1].x += bb[1].x + dd[1].y;
    } else if (V == 2) {   // general, 4 chains: two half blocks interleaved
      float2 bb[4], dd[4];
#pragma unroll
      for (int c = 0; c < 4; c++) { bb[c] = f2(0, 0); dd[c] = f2(0, 0); }
#pragma unroll 4
      for (int m = ROWS / 2 - 1; m >= 0; m--) {
        const float4 q0 = tab[m], q1 = tab[ROWS / 2 + m];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float4 q = (c < 2) ? q0 : q1;
          float2 a_ = __ffma2_rn(w2[c & 1], f2(q.z, q.w), f2(q.x, q.y));
          dd[c] = __ffma2_rn(delta2[c & 1], bb[c], __ffma2_rn(sigma2[c & 1], dd[c], a_));
          bb[c] = __ffma2_rn(sigma2[c & 1], bb[c], dd[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) acc[c].x += bb[c].x + dd[c].y;
    } else if (V == 3) {   // sigma = +1, 4 chains
      float2 bb[4], dd[4];
#pragma unroll
      for (int c = 0; c < 4; c++) { bb[c] = f2(0, 0); dd[c] = f2(0, 0); }
#pragma unroll 4
      for (int m = ROWS / 2 - 1; m >= 0; m--) {
        const float4 q0 = tab[m], q1 = tab[ROWS / 2 + m];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float4 q = (c < 2) ? q0 : q1;
          float2 a_ = __ffma2_rn(w2[c & 1], f2(q.z, q.w), f2(q.x, q.y));
          dd[c] = __ffma2_rn(delta2[c & 1], bb[c], __fadd2_rn(dd[c], a_));
          bb[c] = __fadd2_rn(bb[c], dd[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) acc[c].x += bb[c].x + dd[c].y;
    } else if (V == 4) {   // sigma = +1, lerp folded into the chain: t = fma(w, dY, d) + Y
      float2 bb[2] = {f2(0, 0), f2(0, 0)}, dd[2] = {f2(0, 0), f2(0, 0)};
#pragma unroll 8
      for (int m = ROWS - 1; m >= 0; m--) {
        const float4 q = tab[m];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float2 t = __fadd2_rn(__ffma2_rn(w2[c], f2(q.z, q.w), dd[c]), f2(q.x, q.y));
          dd[c] = __ffma2_rn(delta2[c], bb[c], t);
          bb[c] = __fadd2_rn(bb[c], dd[c]);
        }
      }
      acc[0].x += bb[0].x + dd[0].y; acc[1].x += bb[1].x + dd[1].y;
    } else if (V == 6) {   // standard 3-op, FOUR pairs per lane share {Y, dY}: more operand reuse per load
      float2 b1[4], b2[4];
#pragma unroll
      for (int c = 0; c < 4; c++) { b1[c] = f2(0, 0); b2[c] = f2(0, 0); }
#pragma unroll 8
      for (int m = ROWS - 1; m >= 0; m--) {
        const float4 q = tab[m];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float2 a_ = __ffma2_rn(w4[c], f2(q.z, q.w), f2(q.x, q.y));
          float2 nb = __fadd2_rn(__ffma2_rn(d4[c], b1[c], a_), f2(-b2[c].x, -b2[c].y));
          b2[c] = b1[c]; b1[c] = nb;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) acc[c].x += b1[c].x + b2[c].y;
    } else if (V == 7) {   // sigma = +1 Reinsch, four pairs per lane
      float2 bb[4], dd[4];
#pragma unroll
      for (int c = 0; c < 4; c++) { bb[c] = f2(0, 0); dd[c] = f2(0, 0); }
#pragma unroll 8
      for (int m = ROWS - 1; m >= 0; m--) {
        const float4 q = tab[m];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float2 a_ = __ffma2_rn(w4[c], f2(q.z, q.w), f2(q.x, q.y));
          dd[c] = __ffma2_rn(d4[c], bb[c], __fadd2_rn(dd[c], a_));
          bb[c] = __fadd2_rn(bb[c], dd[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) acc[c].x += bb[c].x + dd[c].y;
    } else if (V == 5) {   // standard (unstable) Clenshaw, 3 ops: b_new = fma(c2, b1, a) - b2
      float2 b1[2] = {f2(0, 0), f2(0, 0)}, b2[2] = {f2(0, 0), f2(0, 0)};
#pragma unroll 8
      for (int m = ROWS - 1; m >= 0; m--) {
        const float4 q = tab[m];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float2 a_ = __ffma2_rn(w2[c], f2(q.z, q.w), f2(q.x, q.y));
          float2 nb = __fadd2_rn(__ffma2_rn(delta2[c], b1[c], a_), f2(-b2[c].x, -b2[c].y));
          b2[c] = b1[c]; b1[c] = nb;
        }
      }
      acc[0].x += b1[0].x + b2[0].y; acc[1].x += b1[1].x + b2[1].y;
    }
  }
  float2 s = f2(acc[0].x + acc[1].x + acc[2].x + acc[3].x, 0);
  if (s.x == 123.456f) out[threadIdx.x] = s;
}

template <typename F> static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  return best;
}
int main() {
  float2 *d; cudaMalloc(&d, 1 << 16);
  int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int blocks = sms * 6 * 4;
  const double clk = khz * 1e3;
  auto rate = [&](double inst_per_thread, float ms) {   // packed instr per cycle per SMSP
    double warps = (double)blocks * 4;
    return warps * inst_per_thread / (ms * 1e-3 * clk) / (sms * 4.0);
  };
  float ms;
  ms = timeit([&] { kA<<<blocks, 128>>>(d, ITERS); });  printf("A  ffma2 1 varying operand      : %.3f inst/clk/SMSP  (%.3f ms)\n", rate(8.0 * ITERS, ms), ms);
  ms = timeit([&] { kB<<<blocks, 128>>>(d, ITERS); });  printf("B  ffma2 3 distinct operands    : %.3f\n", rate(8.0 * ITERS, ms));
  ms = timeit([&] { kC<<<blocks, 128>>>(d, ITERS); });  printf("C  ffma2 shared multiplier      : %.3f\n", rate(8.0 * ITERS, ms));
  ms = timeit([&] { kC2<<<blocks, 128>>>(d, ITERS); }); printf("C2 ffma2 shared mult+addend     : %.3f\n", rate(8.0 * ITERS, ms));
  ms = timeit([&] { kS<<<blocks, 128>>>(d, ITERS); });  printf("S  scalar ffma 3 distinct       : %.3f (scalar inst)\n", rate(16.0 * ITERS, ms));
  const int nblk = 512;
  const double rows = (double)nblk * ROWS;   // pair-rows per chain pair
  ms = timeit([&] { kL<0><<<blocks, 128>>>(d, nblk); }); printf("L0 general 2 chains   : %.3f packed-op/clk/SMSP, %.3f pair-rows/clk/SMSP\n", rate(rows * 8, ms), rate(rows * 2, ms));
  ms = timeit([&] { kL<1><<<blocks, 128>>>(d, nblk); }); printf("L1 sigma=+1 2 chains  : %.3f, %.3f\n", rate(rows * 8, ms), rate(rows * 2, ms));
  ms = timeit([&] { kL<2><<<blocks, 128>>>(d, nblk); }); printf("L2 general 4 chains   : %.3f, %.3f\n", rate(rows * 8, ms), rate(rows * 2, ms));
  ms = timeit([&] { kL<3><<<blocks, 128>>>(d, nblk); }); printf("L3 sigma=+1 4 chains  : %.3f, %.3f\n", rate(rows * 8, ms), rate(rows * 2, ms));
  ms = timeit([&] { kL<4><<<blocks, 128>>>(d, nblk); }); printf("L4 sigma=+1 folded    : %.3f, %.3f\n", rate(rows * 8, ms), rate(rows * 2, ms));
  ms = timeit([&] { kL<5><<<blocks, 128>>>(d, nblk); }); printf("L5 standard 3-op      : %.3f (as 6 ops), %.3f\n", rate(rows * 6, ms), rate(rows * 2, ms));
  ms = timeit([&] { kL<6><<<blocks, 128>>>(d, nblk); }); printf("L6 standard 3-op, 4 pairs/lane : %.3f (as 12 ops), %.3f pair-rows/clk/SMSP\n", rate(rows * 12, ms), rate(rows * 4, ms));
  ms = timeit([&] { kL<7><<<blocks, 128>>>(d, nblk); }); printf("L7 sigma=+1, 4 pairs/lane      : %.3f (as 16 ops), %.3f pair-rows/clk/SMSP\n", rate(rows * 16, ms), rate(rows * 4, ms));
  printf("clock %.0f MHz, %d SMs\n", clk / 1e6, sms);
  return 0;
}
