// FFMA vs FFMA2 issue rate per SM at 8 / 16 / 32 resident warps (register operands only).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[4] = {s, s + 1.f, s + 2.f, s + 3.f}, b[4] = {s * .5f, s * .25f, s * .125f, s * .0625f};
  if (MODE == 0) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(a[i & 3], b[(i >> 2) & 3], acc[i]);
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else {
    unsigned long long acc[16], a2[4], b2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm("mov.b64 %0, {%1, %2};" : "=l"(a2[i]) : "f"(a[i]), "f"(a[3 - i]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(b2[i]) : "f"(b[i]), "f"(b[3 - i]));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"((float)threadIdx.x), "f"((float)i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(a2[i & 3]), "l"(b2[(i >> 2) & 3]));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
      r += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  }
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads = 256; threads <= 1024; threads *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, threads>>>(out, iters, 1.0f); else k<1><<<148, threads>>>(out, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 32.0 * iters * threads;         // per SM (both modes do 32 FMAs per thread per iteration)
      printf("mode %s warps/SM %2d: %.3f ms  %.1f FMA/ns/SM  (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ",
             threads / 32, ms, fma / (ms * 1e6), fma / (ms * 1e6) / 1.965);
    }
  return 0;
}
