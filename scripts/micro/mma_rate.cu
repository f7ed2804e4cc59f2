// tcgen05.mma kind::tf32 issue rate on one SM: cycles per MMA (M 128, K 8) as a function of N, with the A operand in
// shared memory (SS) or in TMEM (TS), and with 1 or 4 CTAs per SM issuing concurrently.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
template <int TS>
__global__ void __launch_bounds__(128) k_rate(int N, int reps, int per_commit, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x;
  for (int i = tid; i < 12288; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3f800000u;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tb)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t t0 = tb;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t da = mk_desc(smem_u32(smem)), db = mk_desc(smem_u32(smem + 16384));
  uint32_t phase = 0;
  long long c0 = clock64();
  if (tid == 0) {
    for (int r = 0; r < reps; r++) {
      for (int i = 0; i < per_commit; i++) {
        if (TS)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                       :: "r"(t0 + 64), "r"(t0 + 8 * (i & 7)), "l"(db), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                       :: "r"(t0 + 64), "l"(da), "l"(db), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      phase ^= 1;
    }
    out[blockIdx.x] = clock64() - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(t0), "n"(128));
}

template <int F16>
__global__ void __launch_bounds__(128) k_multi(int N, int reps, int per_commit, int issuers, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x;
  for (int i = tid; i < 12288; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = F16 ? 0x3c003c00u : 0x3f800000u;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tb)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(issuers));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t t0 = tb;
  const uint32_t idesc = F16 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                             : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
  const uint64_t db = mk_desc(smem_u32(smem + 16384));
  uint32_t phase = 0;
  long long c0 = clock64();
  const int w = tid >> 5;
  for (int r = 0; r < reps; r++) {
    if ((tid & 31) == 0 && w < issuers) {
      for (int i = 0; i < per_commit; i++) {
        if (F16)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                       :: "r"(t0 + 64 + 16 * w), "r"(t0 + 8 * (i & 3)), "l"(db), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                       :: "r"(t0 + 64 + 16 * w), "r"(t0 + 8 * (i & 7)), "l"(db), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1;
    __syncthreads();
  }
  if (tid == 0) out[blockIdx.x] = clock64() - c0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(t0), "n"(128));
}
int main() {
  long long *d, h[1024];
  cudaMalloc(&d, sizeof h);
  cudaFuncSetAttribute(k_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const int Ns[] = {16, 32, 64, 128};
  for (int ctas = 1; ctas <= 4; ctas *= 4)
    for (int ts = 0; ts < 2; ts++)
      for (int pc : {12, 96})
        for (int N : Ns) {
          if (N > 64) continue;         // 64 accumulator columns
          const int reps = 2000 / (pc / 12);
          for (int it = 0; it < 2; it++) {
            if (ts) k_rate<1><<<148 * ctas, 128, 49152>>>(N, reps, pc, d);
            else k_rate<0><<<148 * ctas, 128, 49152>>>(N, reps, pc, d);
            cudaDeviceSynchronize();
          }
          cudaMemcpy(h, d, sizeof(long long) * 148 * ctas, cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < 148 * ctas; i++) s += (double)h[i];
          s /= 148 * ctas;
          printf("ctas/SM %d  %s  N %3d  MMAs/commit %3d: %.1f cycles per MMA per CTA  (%.1f per SM)  err %s\n", ctas, ts ? "TS" : "SS", N, pc,
                 s / ((double)reps * pc), s / ((double)reps * pc) / ctas, cudaGetErrorString(cudaGetLastError()));
        }
  cudaFuncSetAttribute(k_multi<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  cudaFuncSetAttribute(k_multi<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  for (int f16 = 0; f16 < 2; f16++)
    for (int ctas : {1, 4})
      for (int issuers : {1, 2, 4})
        for (int pc : {3, 6, 12}) {
          const int N = 16, reps = 2000;
          for (int it = 0; it < 2; it++) {
            if (f16) k_multi<1><<<148 * ctas, 128, 49152>>>(N, reps, pc, issuers, d);
            else k_multi<0><<<148 * ctas, 128, 49152>>>(N, reps, pc, issuers, d);
            cudaDeviceSynchronize();
          }
          cudaMemcpy(h, d, sizeof(long long) * 148 * ctas, cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < 148 * ctas; i++) s += (double)h[i];
          s /= 148 * ctas;
          printf("%s TS N 16 ctas/SM %d issuers %d MMAs/issuer/commit %2d: %.0f cycles per round (issue + commit + wait + sync)  err %s\n",
                 f16 ? "f16 " : "tf32", ctas, issuers, pc, s / reps, cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
