// mma_rate.cu -- DEVELOPMENT TOOL: raw issue / execution rate of small tcgen05.mma.kind::tf32 instructions
// (M = 128, K = 8, A from TMEM) as a function of N, to size the MMA issue budget of the fused kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include "../../molann_b200/csrc/tc.cuh"
using namespace molann;

__device__ __forceinline__ uint32_t elect1() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

__global__ void rate_kernel(int N, int reps, int ksteps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tptr;
  if (threadIdx.x < 32) {
    const uint32_t leader = elect1();
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t b_a = smem_u32(smem);
    const uint32_t step = 2u * N * 16u, lbo = N * 16u;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int j = 0; j < ksteps; ++j) {
        const uint64_t bd = smem_desc_kmajor(b_a + j * step, lbo, 128);
        if (leader) {
          mma_tf32_ts(tbase + 256, tbase + 64 + 8 * j, bd, idesc, j > 0);
          mma_tf32_ts(tbase + 256, tbase + 8 * j, bd, idesc, 1);
          mma_tf32_ts(tbase + 256, tbase + 8 * j, bd, idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    if (leader) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {16, 32, 64, 128, 256}) {
    const int reps = 64, ks = 8;
    rate_kernel<<<1, 128, 64 * 1024>>>(N, reps, ks, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const int n = reps * ks * 3;
    printf("N=%3d: %d MMAs  issue %.1f cyc/MMA  complete %.1f cyc/MMA  (%s)\n", N, n, (double)h[0] / n, (double)h[1] / n,
           cudaGetErrorString(e));
  }
  return 0;
}
