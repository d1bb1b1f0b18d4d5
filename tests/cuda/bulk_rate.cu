// bulk_rate.cu -- DEVELOPMENT TOOL: throughput of cp.async.bulk (global -> shared, L2-resident source) per SM as a
// function of how many pieces a 64 KB block is split into (pieces are issued back to back on one mbarrier).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bulk_rate bulk_rate.cu && ./bulk_rate
#include <cstdio>
#include "../../molann_b200/csrc/common.cuh"
using namespace molann;

__global__ void rate_kernel(const float* src, int pieces, int reps, int distinct, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = 64 * 1024, piece = total / pieces;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(src) + (distinct ? (size_t)blockIdx.x * total : 0);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      mbar_expect_tx(&bar, total);
      for (int p = 0; p < pieces; ++p) bulk_g2s(smem + p * piece, base + p * piece, piece, &bar);
      mbar_wait(&bar, r & 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  float* d; cudaMalloc(&d, 148 * 64 * 1024);
  cudaMemset(d, 0, 148 * 64 * 1024);
  long long* o; cudaMalloc(&o, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int distinct = 0; distinct < 2; ++distinct)
    for (int pieces : {1, 2, 4, 8, 16, 64}) {
      const int reps = 200;
      rate_kernel<<<148, 32, 64 * 1024>>>(d, pieces, reps, distinct, o);
      cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
      printf("%s source, 64 KB in %2d pieces: %.0f cycles per block  (%.1f B/clk/SM)\n", distinct ? "per-SM " : "shared ", pieces,
             (double)h / reps, 65536.0 * reps / h);
    }
  return 0;
}
