// K1 on the 5th-generation tensor cores (tcgen05 + TMEM): the harmonic loop of generateHarmonics
// (R/source.R:389-419) as a contraction.
//
// The rows of an epoch are integer multiples j of theta' = 2 pi integr / (nSubharm + 1).  With j = KR b + m,
//   sum_j a_j(u) sin(j theta'_u) = Re sum_b e^{i KR b theta'_u} [ S_b(u) - i C_b(u) ],
//   C_b(u) = sum_m a_{KR b + m}(u) cos(m theta'_u),   S_b(u) = sum_m a_{KR b + m}(u) sin(m theta'_u),
// and a_j(u) = Y_j + w(u) dY_j inside one interval of approx() (source.R:403-405), so for the samples of one
// interval C and S are small GEMMs  [128 samples x KR] . [KR x 2 blocks]  of a trig matrix (KR values per sample,
// one complex rotation each) against the interval's amplitude column -- instead of one recurrence step per
// (row, sample) on the FMA pipe.  A CTA (128 threads) takes one interval of approx() (one glottal cycle of one
// epoch, a "unit") at a time:
//   * it stages the amplitude operand of the interval ONCE: Y | dY of up to 1024 rows, split hi + lo in TF32,
//     from K3's FP32 table into shared memory in the canonical K-major core-matrix layout;
//   * per tile of 128 samples (thread = sample = TMEM lane) every thread writes its trig row cos / sin(m theta'),
//     m < KR, hi + lo; one thread issues 12 tcgen05.mma kind::tf32 (M 128, N = 2 x blocks <= 64, K 8) per pass of
//     512 rows: 2 k-steps x 3xTF32 terms (hi hi + lo hi + hi lo: FP32-grade products) x (cos, sin), accumulators
//     in TMEM (128 columns per CTA, four CTAs per SM);
//   * tcgen05.ld hands every thread its own lane: C_b, S_b (Y and dY parts) of the pass for ITS sample, reduced
//     in registers by a complex Horner recurrence in e^{i KR theta'}; the pass offset e^{i 512 p theta'} comes
//     from the FP64 phase.
// Phase: as in the FMA kernel, FP64 closed form per spline piece (K0's quartic), re-anchored at every knot.
// Numerics: scripts/micro/k1_umma.cu; the GPU parity tests run through this kernel.
#include "engine.cuh"
#include <cstdlib>

#define TC_KR 16                         // rows per block = K of the contraction
#define TC_NB 32                         // blocks per chunk
#define TC_NCOL (2 * TC_NB)              // N of one MMA: Y and dY of every block
#define TC_CHUNK (TC_KR * TC_NB)         // 512 rows per chunk
#define TC_A_BYTES (128 * TC_KR * 4)     // one trig operand: 8 KB
#define TC_B_BYTES (TC_NCOL * TC_KR * 4) // one amplitude operand: 4 KB
#define TC_TILE 128

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tc_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
// K-major, no swizzle, in 16-byte units ((8, n), 2) : ((1, SBO), LBO): a core matrix is 8 rows x 16 bytes;
// the next 16 bytes of K are LBO = 128 B further, the next 8 rows SBO = (KR / 4) * 128 B further
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(((TC_KR / 4) * 128) >> 4) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version of sm_100
  return d;
}
__device__ __forceinline__ int tc_off(int r, int k) { return (r >> 3) * ((TC_KR / 4) * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                 "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}

__device__ int g_tc_timeout = 0;          // an MMA completion barrier that never flipped (reported by the host)
int synth_tc_timeout_flag() { int v = 0; cudaMemcpyFromSymbol(&v, g_tc_timeout, sizeof v); return v; }

// tcgen05.ld 32x32b.x16: 16 consecutive columns of this thread's lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
// x = hi + lo with hi on the TF32 grid (round to nearest, ties away) and lo exact in FP32; the tensor core reads
// the top 19 bits of lo, i.e. 2^-22 of x is kept -- three full-rate integer / FP32 operations instead of two cvt
__device__ __forceinline__ void tc_split(float x, uint32_t &hi, uint32_t &lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// Work list: one unit per interval of approx() of every epoch.  One thread per syllable; intervals without
// samples get kbeg == kend.
__global__ void k_build_units_tc(const sgb_syllable *syl, const SylCtrl *ctrl, int S, const SylLayout *lay, const Pools P, TcUnit *units) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const SylCtrl &C = ctrl[s];
  if (C.status != SGB_OK || C.nGC == 0) return;
  const int64_t o = P.gc_off[s];
  const int32_t *gcup = P.gcup + o;
  const double *kt = P.kt + o;
  const int G = C.nGC;
  int t = lay[s].pad;                     // first unit of this syllable
  const int t_end = t + C.tiles_tc;
  int a = 0;
  for (int e = 0; e < C.nEpochs; e++) {
    const int g_first = C.ep_start[e] - 1, g_last = C.ep_end[e] - 1;
    const int x_first_i = gcup[g_first];
    const double x_first = (double)x_first_i, x_last = (double)gcup[g_last];
    const int Ne = gcup[C.ep_end[e]] - x_first_i + 1;
    const double by = (x_last - x_first) / (double)(Ne - 1);
    const int nknots = g_last - g_first + 1;
    const int nsub = C.vf_active ? C.ep_nsub[e] : 0;
    const int J = C.ep_rows[e];
    auto vfun = [&](int k) { return (k >= Ne - 1) ? x_last : (x_first + (double)k * by); };
    int kcur = 0;
    for (int i = 0; i <= nknots - 2; i++) {
      int kend = Ne;
      if (i < nknots - 2) {               // first sample whose coordinate reaches the next knot
        const double X = (double)gcup[g_first + i + 1];
        int k = (int)ceil((X - x_first) / by);
        k = max(kcur, min(k, Ne));
        while (k > kcur && vfun(k - 1) >= X) k--;
        while (k < Ne && vfun(k) < X) k++;
        kend = k;
      }
      const double u = (double)(x_first_i + kcur);
      while (a < G - 1 && u >= kt[a + 1]) a++;
      TcUnit U;
      U.col_off = lay[s].amp_off + C.ep_amp_off[e] + (int64_t)i * J;
      U.wave_off = lay[s].wave_off + C.ep_wave_off[e];
      U.pc_off = o;
      U.x_first = x_first; U.by = by; U.x_last = x_last;
      U.inv_sr_np1 = 1.0 / (syl[s].samplingRate * (double)(nsub + 1));
      U.Ne = Ne; U.kbeg = kcur; U.kend = kend; U.xg = gcup[g_first + i]; U.xn = gcup[g_first + i + 1];
      U.a_lo = a; U.G = G; U.J = J; U.epmax_idx = s * SGB_MAX_EPOCHS + e; U.pad = 0;
      if (t < t_end) units[t++] = U;
      kcur = kend;
    }
    while (a > 0 && (double)gcup[g_last] < kt[a]) a--;     // epochs overlap by one cycle
  }
  TcUnit Z = {};
  while (t < t_end) units[t++] = Z;
}

#define TC_PASS_ROWS 512                          // rows of one pass: 32 blocks of KR
#define TC_B_PASSES 2                             // passes whose amplitude operand is resident (1024 rows)
#define TC_BP_BYTES (64 * TC_KR * 4)              // one amplitude image of a pass: 64 columns (Y, dY of 32 blocks) x KR
#define TC_SMEM (4 * TC_A_BYTES + TC_B_PASSES * 2 * TC_BP_BYTES)

__global__ void __launch_bounds__(128, 4)
k_synth_tc(const TcUnit *__restrict__ units, int nunits, const double *__restrict__ pc, const float4 *__restrict__ amp,
           float *__restrict__ wave, int *__restrict__ epmax) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                       // cos_hi, cos_lo, sin_hi, sin_lo
  uint8_t *sB = smem + 4 * TC_A_BYTES;      // per resident pass: hi, lo
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&tmem_base_s)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t aB = tc_smem_u32(sA), bB = tc_smem_u32(sB), barA = tc_smem_u32(&bar);
  uint32_t phase = 0u;
  bool dead = false;

  for (int ui = blockIdx.x; ui < nunits && !dead; ui += gridDim.x) {
    const TcUnit U = units[ui];
    if (U.kend <= U.kbeg) continue;
    const int J = U.J;
    const float4 *__restrict__ col = amp + U.col_off;               // {Y, Y, dY, dY} of the interval, row j - 1
    const double *__restrict__ pcs = pc + SYNTH_PC * U.pc_off;
    const int x_first_i = (int)U.x_first;
    const float inv_dx = __frcp_rn((float)(U.xn - U.xg));
    const int nrows = J + 1;                                        // rows j = 0 (zero) .. J
    const int nsuper = (nrows + TC_B_PASSES * TC_PASS_ROWS - 1) / (TC_B_PASSES * TC_PASS_ROWS);

    for (int sp = 0; sp < nsuper; sp++) {
      const int row0 = sp * TC_B_PASSES * TC_PASS_ROWS;
      const int rows_here = min(nrows - row0, TC_B_PASSES * TC_PASS_ROWS);
      const int blocks_here = ((rows_here + TC_KR - 1) / TC_KR + 7) & ~7;    // N = 2 x blocks is a multiple of 16
      // ---- amplitude operand: rows row0 .. row0 + 16 blocks_here - 1 ----
      for (int r = tid; r < blocks_here * TC_KR; r += 128) {
        const int j = row0 + r;
        float y = 0.f, dy = 0.f;
        if (j >= 1 && j <= J) { const float4 q = __ldg(&col[j - 1]); y = q.x; dy = q.z; }
        const int p = r / TC_PASS_ROWS, rr = r % TC_PASS_ROWS, b = rr / TC_KR, m = rr % TC_KR;
        uint32_t yh, yl, dh, dl;
        tc_split(y, yh, yl); tc_split(dy, dh, dl);
        uint8_t *hi = sB + p * (2 * TC_BP_BYTES) + tc_off(2 * b, m), *lo = hi + TC_BP_BYTES;
        *reinterpret_cast<uint32_t *>(hi) = yh; *reinterpret_cast<uint32_t *>(hi + 16) = dh;
        *reinterpret_cast<uint32_t *>(lo) = yl; *reinterpret_cast<uint32_t *>(lo + 16) = dl;
      }
      const int npass = (blocks_here * TC_KR + TC_PASS_ROWS - 1) / TC_PASS_ROWS;

      for (int k0 = U.kbeg; k0 < U.kend && !dead; k0 += TC_TILE) {
        // ---- this thread's sample: interval weight and phase ----
        const bool live = k0 + tid < U.kend;
        const int k = live ? k0 + tid : k0;
        const double v = (k >= U.Ne - 1) ? U.x_last : (U.x_first + (double)k * U.by);
        const float ww = (float)(v - (double)U.xg) * inv_dx;
        const double u = (double)(x_first_i + k);
        int a = U.a_lo;
        while (a < U.G - 1 && u >= pcs[SYNTH_PC * (a + 1)]) a++;
        const double *pp = pcs + SYNTH_PC * a;
        const double M = u - pp[0];
        const double ph = fma(fma(fma(fma(pp[5], M, pp[4]), M, pp[3]), M, pp[2]), M, pp[1]);
        double x = ph * U.inv_sr_np1;             // integr / (nSubharm + 1), in cycles
        x -= floor(x);
        // ---- trig operand rows: cos / sin (m theta'), m < KR, hi + lo ----
        {
          float s1, c1;
          sincospif(2.0f * (float)x, &s1, &c1);
          float cm = 1.f, sm = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < TC_KR / 4; k4++) {
            uint32_t ch[4], cl[4], sh[4], sl[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
              tc_split(cm, ch[q], cl[q]); tc_split(sm, sh[q], sl[q]);
              const float c = cm * c1 - sm * s1, sn = cm * s1 + sm * c1; cm = c; sm = sn;
            }
            const int of = tc_off(tid, 4 * k4);
            *reinterpret_cast<uint4 *>(sA + 0 * TC_A_BYTES + of) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
            *reinterpret_cast<uint4 *>(sA + 1 * TC_A_BYTES + of) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
            *reinterpret_cast<uint4 *>(sA + 2 * TC_A_BYTES + of) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
            *reinterpret_cast<uint4 *>(sA + 3 * TC_A_BYTES + of) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
          }
        }
        float zc, zs;                            // z = e^{i KR theta'}
        { double q = x * (double)TC_KR; q -= floor(q); sincospif(2.0f * (float)q, &zs, &zc); }
        float acc = (sp > 0 && live) ? wave[U.wave_off + k] : 0.f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");

        for (int p = 0; p < npass; p++) {
          const int nb = min(blocks_here - p * (TC_PASS_ROWS / TC_KR), TC_PASS_ROWS / TC_KR);   // blocks of this pass: 8 .. 32
          if (tid == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * nb) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t bP = bB + p * (2 * TC_BP_BYTES);
#pragma unroll
            for (int ks = 0; ks < 2; ks++)
#pragma unroll
              for (int t3 = 0; t3 < 3; t3++) {
                const int ah = (t3 == 1) ? 1 : 0, bh = (t3 == 2) ? 1 : 0;
                const uint32_t acc_flag = (ks | t3) ? 1u : 0u;
                const uint64_t bd = tc_desc(bP + bh * TC_BP_BYTES + ks * 256);
                tc_mma(tmem_base, tc_desc(aB + (0 + ah) * TC_A_BYTES + ks * 256), bd, idesc, acc_flag);
                tc_mma(tmem_base + 64, tc_desc(aB + (2 + ah) * TC_A_BYTES + ks * 256), bd, idesc, acc_flag);
              }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(barA) : "memory");
          }
          uint32_t ok = 0;
          unsigned long long polls = 0;
          while (!ok) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(barA), "r"(phase) : "memory");
            if (!ok && ++polls > (1ull << 24)) { g_tc_timeout = 1; dead = true; break; }
          }
          phase ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;");
          // ---- R = sum_b z^b (S_b - i C_b), Horner from the top block ----
          float Rr = 0.f, Ri = 0.f;
          int b0 = nb;
          if (b0 & 8) {                          // tail group of 8 blocks
            b0 -= 8;
            float Cv[16], Sv[16];
            tc_ld16(lane_addr + 2 * b0, Cv); tc_ld16(lane_addr + 64 + 2 * b0, Sv);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int b = 7; b >= 0; b--) {
              const float Cc = fmaf(ww, Cv[2 * b + 1], Cv[2 * b]), Ss = fmaf(ww, Sv[2 * b + 1], Sv[2 * b]);
              const float nr = fmaf(Rr, zc, fmaf(-Ri, zs, Ss)), ni = fmaf(Rr, zs, fmaf(Ri, zc, -Cc));
              Rr = nr; Ri = ni;
            }
          }
          while (b0 > 0) {
            b0 -= 16;
            float Cv[32], Sv[32];
            tc_ld32(lane_addr + 2 * b0, Cv); tc_ld32(lane_addr + 64 + 2 * b0, Sv);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int b = 15; b >= 0; b--) {
              const float Cc = fmaf(ww, Cv[2 * b + 1], Cv[2 * b]), Ss = fmaf(ww, Sv[2 * b + 1], Sv[2 * b]);
              const float nr = fmaf(Rr, zc, fmaf(-Ri, zs, Ss)), ni = fmaf(Rr, zs, fmaf(Ri, zc, -Cc));
              Rr = nr; Ri = ni;
            }
          }
          const int jb = row0 + p * TC_PASS_ROWS;             // first row of the pass
          if (jb == 0) acc += Rr;
          else {
            float rs, rc;
            double q = x * (double)jb; q -= floor(q);
            sincospif(2.0f * (float)q, &rs, &rc);
            acc = fmaf(rc, Rr, fmaf(-rs, Ri, acc));
          }
          asm volatile("tcgen05.fence::before_thread_sync;");
          __syncthreads();                                    // TMEM (and, after the last pass, the trig operand) is free
          asm volatile("tcgen05.fence::after_thread_sync;");
        }
        if (live) wave[U.wave_off + k] = acc;
        if (sp == nsuper - 1) {      // max |w| of the epoch (tolerance of the zero-crossing searches in K6)
          float m = live ? fabsf(acc) : 0.0f;
#pragma unroll
          for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));
          if (lane == 0 && m > 0.0f) atomicMax(&epmax[U.epmax_idx], float_to_ordered(m));
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(128));
}

void launch_build_tiles_tc(const sgb_syllable *syl, const SylCtrl *ctrl, int S, const SylLayout *lay, const Pools &P, TcUnit *units, cudaStream_t st) {
  if (S <= 0) return;
  k_build_units_tc<<<(S + 127) / 128, 128, 0, st>>>(syl, ctrl, S, lay, P, units);
}

cudaError_t launch_synth_tc(const TcUnit *units, int n_units, const Pools &P, const float4 *amp, float *wave, int *epmax, cudaStream_t st) {
  if (n_units <= 0) return cudaSuccess;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_synth_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(n_units, 4 * sms);               // persistent: four CTAs (4 x 128 TMEM columns) per SM
  k_synth_tc<<<grid, 128, TC_SMEM, st>>>(units, n_units, P.pc, amp, wave, epmax);
  return cudaGetLastError();
}
