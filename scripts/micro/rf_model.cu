// micro-benchmark: cost of FP32x2 / FP32 FMA-pipe instructions on sm_100a as a function of how many
// FRESH register operands each reads (operands repeated from the previous instruction can come from
// the operand-reuse cache).  All operands are per-lane values loaded from memory (nothing folds).
// Prints cycles per warp instruction per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define NCH 8
typedef unsigned long long u64;

#define F2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define A2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define M2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define F1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define A1(d, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))

template <int T>
__global__ void __launch_bounds__(128, 6) k2(const u64 *in, u64 *out, int it) {
  u64 d[NCH], b[NCH], c[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    d[i] = in[threadIdx.x + 128 * i]; b[i] = in[threadIdx.x + 128 * (i + 8)]; c[i] = in[threadIdx.x + 128 * (i + 16)];
  }
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (T == 1) F2(d[i], d[i], b[0], c[0]);          // 1 fresh (acc in slot A)
      if (T == 2) F2(d[i], d[i], b[i], c[0]);          // 2 fresh
      if (T == 3) F2(d[i], d[i], b[i], c[i]);          // 3 fresh
      if (T == 4) F2(d[i], b[i], c[i], d[i]);          // 3 fresh (acc in slot C)
      if (T == 5) F2(d[i], b[0], c[0], d[i]);          // 1 fresh (acc in slot C)
      if (T == 6) F2(d[i], b[0], c[i], d[i]);          // 2 fresh (A repeated)
      if (T == 7) A2(d[i], d[i], b[i]);                // add, 2 fresh
      if (T == 8) A2(d[i], d[i], b[0]);                // add, 1 fresh
      if (T == 9) M2(d[i], d[i], b[i]);                // mul, 2 fresh
      if (T == 10) F2(d[i], d[i], d[i], c[i]);         // 2 distinct, one used twice
      if (T == 11) F2(d[i], b[0], d[i], c[i]);         // 2 fresh (acc in slot B)
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s ^= d[i];
  if (s == 0x123456789abcull) out[threadIdx.x] = s;
}
template <int T>
__global__ void __launch_bounds__(128, 6) k1(const float *in, float *out, int it) {
  float d[2 * NCH], b[2 * NCH], c[2 * NCH];
#pragma unroll
  for (int i = 0; i < 2 * NCH; i++) {
    d[i] = in[threadIdx.x + 128 * i]; b[i] = in[threadIdx.x + 128 * (i + 16)]; c[i] = in[threadIdx.x + 128 * (i + 32)];
  }
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int i = 0; i < 2 * NCH; i++) {
      if (T == 1) F1(d[i], d[i], b[0], c[0]);
      if (T == 2) F1(d[i], d[i], b[i], c[0]);
      if (T == 3) F1(d[i], d[i], b[i], c[i]);
      if (T == 4) F1(d[i], b[i], c[i], d[i]);
      if (T == 5) F1(d[i], b[0], c[0], d[i]);
      if (T == 6) F1(d[i], b[0], c[i], d[i]);
      if (T == 7) A1(d[i], d[i], b[i]);
      if (T == 8) A1(d[i], d[i], b[0]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * NCH; i++) s += d[i];
  if (s == 123.456f) out[threadIdx.x] = s;
}

static float timeit(void (*f)(void)) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  return best;
}
static u64 *g_in; static u64 *g_out; static int g_blocks;
template <int T> static void run2() { k2<T><<<g_blocks, 128>>>(g_in, g_out, ITERS); }
template <int T> static void run1() { k1<T><<<g_blocks, 128>>>((const float *)g_in, (float *)g_out, ITERS); }
int main() {
  int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  cudaMalloc(&g_in, 1 << 20); cudaMalloc(&g_out, 1 << 16);
  float *h = (float *)malloc(1 << 20);
  for (int i = 0; i < (1 << 18); i++) h[i] = 0.5f + 1e-6f * (i % 1000);
  cudaMemcpy(g_in, h, 1 << 20, cudaMemcpyHostToDevice);
  g_blocks = sms * 6 * 4;
  const double clk = khz * 1e3;
  auto cyc = [&](double inst_per_thread, float ms) {
    double warps_per_smsp = (double)g_blocks * 4 / (sms * 4.0);
    return ms * 1e-3 * clk / (warps_per_smsp * inst_per_thread);
  };
  const char *n2[] = {"", "ffma2 1 fresh (acc A)", "ffma2 2 fresh", "ffma2 3 fresh", "ffma2 3 fresh (acc C)", "ffma2 1 fresh (acc C)",
                      "ffma2 2 fresh (A repeated)", "fadd2 2 fresh", "fadd2 1 fresh", "fmul2 2 fresh", "ffma2 d,d,c", "ffma2 2 fresh (acc B)"};
  void (*f2[])(void) = {nullptr, run2<1>, run2<2>, run2<3>, run2<4>, run2<5>, run2<6>, run2<7>, run2<8>, run2<9>, run2<10>, run2<11>};
  for (int t = 1; t <= 11; t++) printf("%-28s : %.2f cycles/inst/SMSP\n", n2[t], cyc((double)NCH * ITERS, timeit(f2[t])));
  const char *n1[] = {"", "ffma 1 fresh (acc A)", "ffma 2 fresh", "ffma 3 fresh", "ffma 3 fresh (acc C)", "ffma 1 fresh (acc C)",
                      "ffma 2 fresh (A repeated)", "fadd 2 fresh", "fadd 1 fresh"};
  void (*f1[])(void) = {nullptr, run1<1>, run1<2>, run1<3>, run1<4>, run1<5>, run1<6>, run1<7>, run1<8>};
  for (int t = 1; t <= 8; t++) printf("%-28s : %.2f cycles/inst/SMSP\n", n1[t], cyc(2.0 * NCH * ITERS, timeit(f1[t])));
  return 0;
}
