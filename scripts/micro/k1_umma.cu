// Microbenchmark, second half of VERDICT r01 item 7: the same contraction as k1_mma.cu (rows j = KR b + m,
// C_b(u) = sum_m a_{KR b+m}(u) cos(m th_u), S_b likewise, then sum_b [sin(KR b th) C_b + cos(KR b th) S_b])
// on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the 3xTF32 split, operands staged in
// shared memory in the canonical K-major core-matrix layout, accumulators in TMEM, epilogue through
// tcgen05.ld (one sample per thread = one TMEM lane, so the sum over blocks is thread-local).
//   M = 128 samples of one tile, N = 2 NB (Y and dY of NB blocks), K = 8 per instruction, KR = 16.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o k1_umma k1_umma.cu
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int KR = 16, NB = 32, NCOL = 2 * NB;         // rows per block, blocks per chunk, N of one MMA
constexpr int CHUNK_ROWS = KR * NB;                     // 512 rows per chunk
constexpr int A_BYTES = 128 * KR * 4;                   // one trig matrix: 8 KB
constexpr int B_BYTES = NCOL * KR * 4;                  // one amplitude matrix: 4 KB
constexpr double TWO_PI = 6.283185307179586;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tf32_hi(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
// K-major, no swizzle: ((8, n), 2) : ((16 B, SBO), LBO); LBO = 128 B (next core matrix along K), SBO = (KR / 4) * 128 B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;                      // leading byte offset
  d |= (uint64_t)(((KR / 4) * 128) >> 4) << 32;         // stride byte offset
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  return d;                                             // base offset 0, layout type 0 = no swizzle
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                 "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
// byte offset of element (row r, k) inside one operand matrix
__device__ __host__ __forceinline__ int op_off(int r, int k) { return (r >> 3) * ((KR / 4) * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4; }

__device__ int g_fail = 0;

// tabB: per (column, chunk) the smem image of the two amplitude operands: [hi 4 KB | lo 4 KB]
__global__ void __launch_bounds__(128) k_umma(const uint8_t *__restrict__ tabB, int nchunks, const double *__restrict__ th,
                                              const float *__restrict__ w, const int *__restrict__ col, int ntiles, float *__restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                       // cos_hi, cos_lo, sin_hi, sin_lo: 4 x 8 KB
  uint8_t *sB = smem + 4 * A_BYTES;         // 2 buffers x (hi, lo): 2 x 8 KB
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase[2] = {0u, 0u};

  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint8_t *Bg = tabB + (size_t)col[t] * nchunks * (2 * B_BYTES);
    // ---- trig operand rows of this thread's sample: cos / sin (m th), m = 0..KR-1, split hi / lo ----
    const double thu = th[(size_t)t * 128 + tid];
    const double fr = thu - floor(thu);
    float s1, c1;
    sincospif(2.0f * (float)fr, &s1, &c1);
    {
      float cm = 1.f, sm = 0.f;
      uint32_t ch[KR], cl[KR], sh[KR], sl[KR];
#pragma unroll
      for (int m = 0; m < KR; m++) {
        ch[m] = tf32_hi(cm); cl[m] = tf32_hi(cm - __uint_as_float(ch[m]));
        sh[m] = tf32_hi(sm); sl[m] = tf32_hi(sm - __uint_as_float(sh[m]));
        const float c = cm * c1 - sm * s1, s = cm * s1 + sm * c1; cm = c; sm = s;
      }
#pragma unroll
      for (int k4 = 0; k4 < KR / 4; k4++) {
        const int o = op_off(tid, 4 * k4);
        *reinterpret_cast<uint4 *>(sA + 0 * A_BYTES + o) = make_uint4(ch[4 * k4], ch[4 * k4 + 1], ch[4 * k4 + 2], ch[4 * k4 + 3]);
        *reinterpret_cast<uint4 *>(sA + 1 * A_BYTES + o) = make_uint4(cl[4 * k4], cl[4 * k4 + 1], cl[4 * k4 + 2], cl[4 * k4 + 3]);
        *reinterpret_cast<uint4 *>(sA + 2 * A_BYTES + o) = make_uint4(sh[4 * k4], sh[4 * k4 + 1], sh[4 * k4 + 2], sh[4 * k4 + 3]);
        *reinterpret_cast<uint4 *>(sA + 3 * A_BYTES + o) = make_uint4(sl[4 * k4], sl[4 * k4 + 1], sl[4 * k4 + 2], sl[4 * k4 + 3]);
      }
    }
    // rotators: block step e^{i KR th}, start at block 0
    const float ww = w[(size_t)t * 128 + tid];
    float stc, sts;
    { double a = thu * (double)KR; a -= floor(a); sincospif(2.0f * (float)a, &sts, &stc); }
    float rc = 1.f, rs = 0.f, acc = 0.f;

    auto load_B = [&](int c, int buf) {      // 8 KB per chunk: 128 threads x 4 x 16 B
      const uint4 *src = reinterpret_cast<const uint4 *>(Bg + (size_t)c * (2 * B_BYTES));
      uint4 *dst = reinterpret_cast<uint4 *>(sB + buf * (2 * B_BYTES));
#pragma unroll
      for (int i = 0; i < 4; i++) dst[tid + 128 * i] = __ldg(&src[tid + 128 * i]);
    };
    auto issue = [&](int buf) {              // 12 MMAs: 2 k-steps x 3 split terms x (cos, sin)
      const uint32_t aB = smem_u32(sA), bB = smem_u32(sB + buf * (2 * B_BYTES));
      const uint32_t dC = tmem_base + buf * 128, dS = dC + NCOL;
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
#pragma unroll
        for (int sp = 0; sp < 3; sp++) {
          const int ah = (sp == 1) ? 1 : 0, bh = (sp == 2) ? 1 : 0;
          const uint32_t acc_flag = (ks | sp) ? 1u : 0u;
          const uint64_t bd = make_desc(bB + bh * B_BYTES + ks * 256);
          umma_tf32(dC, make_desc(aB + (0 + ah) * A_BYTES + ks * 256), bd, idesc, acc_flag);
          umma_tf32(dS, make_desc(aB + (2 + ah) * A_BYTES + ks * 256), bd, idesc, acc_flag);
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[buf])) : "memory");
    };
    load_B(0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) issue(0);
    for (int c = 0; c < nchunks; c++) {
      const int buf = c & 1;
      if (c + 1 < nchunks) {                 // stage and issue the next chunk while this one is being reduced
        load_B(c + 1, buf ^ 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) issue(buf ^ 1);
      }
      // wait for this chunk's MMAs
      uint32_t ok = 0; unsigned long long polls = 0;
      while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar[buf])), "r"(phase[buf]) : "memory");
        if (!ok && ++polls > (1ull << 24)) { g_fail = 1; break; }
      }
      phase[buf] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 128;
      float Cy[32], Cd[32], Sy[32], Sd[32];
      tmem_ld32(lane_addr, Cy); tmem_ld32(lane_addr + 32, Cd); tmem_ld32(lane_addr + 64, Sy); tmem_ld32(lane_addr + 96, Sd);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int b = 0; b < NB; b++) {
        const float C = fmaf(ww, Cd[b], Cy[b]), S = fmaf(ww, Sd[b], Sy[b]);
        acc = fmaf(rs, C, fmaf(rc, S, acc));
        const float cn = rc * stc - rs * sts, sn = rc * sts + rs * stc; rc = cn; rs = sn;
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();                       // everyone is done with TMEM buffer `buf` and smem buffer `buf`
    }
    out[(size_t)t * 128 + tid] = acc;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(256));
}

__global__ void k_ref(const double *Y, const double *dY, int J, int Jpad, const double *th, const float *w, const int *col,
                      int nsamp, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsamp) return;
  const double *y = Y + (size_t)col[i / 128] * Jpad, *d = dY + (size_t)col[i / 128] * Jpad;
  double s = 0, ww = w[i], t = th[i];
  for (int j = 1; j < J; j++) s += (y[j] + ww * d[j]) * sin(TWO_PI * (double)j * t);
  out[i] = s;
}

int main(int argc, char **argv) {
  const int J = argc > 1 ? atoi(argv[1]) : 1024;
  const int ntiles = argc > 2 ? atoi(argv[2]) : 1 << 16;     // 128-sample tiles
  const int ncol = 1024, nchunks = (J + CHUNK_ROWS - 1) / CHUNK_ROWS, Jpad = nchunks * CHUNK_ROWS;
  std::vector<double> Y((size_t)ncol * Jpad, 0.0), dY((size_t)ncol * Jpad, 0.0), th((size_t)ntiles * 128);
  std::vector<uint8_t> tabB((size_t)ncol * nchunks * 2 * B_BYTES);
  std::vector<float> w((size_t)ntiles * 128);
  std::vector<int> col(ntiles);
  srand(1);
  auto rnd = [] { return rand() / (double)RAND_MAX; };
  auto sp = [](double x, float &hi, float &lo) { float f = (float)x; uint32_t b; memcpy(&b, &f, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&hi, &b, 4); float rr = f - hi; memcpy(&b, &rr, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&lo, &b, 4); };
  for (int c = 0; c < ncol; c++) {
    const double slope = -1.0 - 5.0 * rnd();
    for (int j = 1; j < J; j++) {
      double a = pow(2.0, slope * log2((double)j) / 10.0) * (0.7 + 0.6 * rnd());
      Y[(size_t)c * Jpad + j] = a; dY[(size_t)c * Jpad + j] = a * 0.2 * (rnd() - 0.5);
    }
    for (int ch = 0; ch < nchunks; ch++) {
      uint8_t *img = tabB.data() + ((size_t)c * nchunks + ch) * 2 * B_BYTES;
      for (int n = 0; n < NCOL; n++)
        for (int m = 0; m < KR; m++) {
          const int j = ch * CHUNK_ROWS + KR * (n % NB) + m;
          const double v = (n < NB) ? Y[(size_t)c * Jpad + j] : dY[(size_t)c * Jpad + j];
          float hi, lo; sp(v, hi, lo);
          memcpy(img + op_off(n, m), &hi, 4); memcpy(img + B_BYTES + op_off(n, m), &lo, 4);
        }
    }
  }
  for (int t = 0; t < ntiles; t++) {
    col[t] = rand() % ncol;
    double f0 = 50 + 70 * rnd(), ph = 1000 * rnd();
    for (int u = 0; u < 128; u++) { th[(size_t)t * 128 + u] = (ph + u * f0 / 48000.0) / 5.0; w[(size_t)t * 128 + u] = (float)rnd(); }
  }
  uint8_t *dtab; double *dth, *dY64, *ddY64, *dref; float *dw, *dout; int *dcol;
  CK(cudaMalloc(&dtab, tabB.size())); CK(cudaMalloc(&dth, th.size() * 8)); CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMalloc(&dcol, col.size() * 4)); CK(cudaMalloc(&dout, w.size() * 4));
  CK(cudaMemcpy(dtab, tabB.data(), tabB.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dth, th.data(), th.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcol, col.data(), col.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 4 * A_BYTES + 2 * 2 * B_BYTES;      // 48 KB
  CK(cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = 148 * 2;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; i++) k_umma<<<grid, 128, smem>>>(dtab, nchunks, dth, dw, dcol, ntiles, dout);
  CK(cudaDeviceSynchronize());
  int fail = 0; CK(cudaMemcpyFromSymbol(&fail, g_fail, 4));
  if (fail) { printf("an MMA completion barrier timed out\n"); return 1; }
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int i = 0; i < reps; i++) k_umma<<<grid, 128, smem>>>(dtab, nchunks, dth, dw, dcol, ntiles, dout);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  const double ps = (double)ntiles * 128 * (J - 1);
  printf("J %d (padded %d), %d tiles of 128: %.3f ms, %.3e partial-samples/s, %.2f TFLOP/s at 6 flop per partial-sample\n", J, Jpad, ntiles, ms, ps / (ms * 1e-3), 6 * ps / (ms * 1e-3) / 1e12);
  const int nchk = 512 * 128;
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
