// micro-benchmark: legacy mma.sync throughput on sm_100a (fp16 m16n8k16 and tf32 m16n8k8, fp32 accumulate)
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__global__ void k_h(float *out, int iters) {
  unsigned a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
  float c[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.f) out[0] = s;
}
__global__ void k_t(float *out, int iters) {
  unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
  float c[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.f) out[0] = s;
}
int main() {
  float *d; cudaMalloc(&d, 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int warps = 4; warps <= 16; warps *= 2) {
    int iters = 20000, blocks = sms * 2, threads = warps * 32;
    for (int kind = 0; kind < 2; kind++) {
      float best = 1e9;
      for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        if (kind == 0) k_h<<<blocks, threads>>>(d, iters); else k_t<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double flop = (double)blocks * warps * iters * 8.0 * (kind == 0 ? 2.0 * 16 * 8 * 16 : 2.0 * 16 * 8 * 8);
      printf("%s warps/CTA %d (2 CTA/SM): %.1f TFLOP/s\n", kind == 0 ? "mma.sync f16 m16n8k16" : "mma.sync tf32 m16n8k8", warps, flop / best / 1e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
