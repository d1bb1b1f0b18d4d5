// tc_probe.cu -- stand-alone GPU probe of the tcgen05 building blocks in molann_b200/csrc/tc.cuh:
//   D[128,N] = A[128,K] * W[N,K]^T with 3xTF32, A staged through TMEM (TS) or shared memory (SS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tests/cuda/tc_probe.cu ; run on a B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../molann_b200/csrc/tc.cuh"

using namespace molann;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                    float* __restrict__ D, int K, int N, int mode, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 16);
  unsigned char* Bhi = smem + 1024;
  unsigned char* Blo = Bhi + K * N * 4;
  unsigned char* Ahi = Blo + K * N * 4;
  unsigned char* Alo = Ahi + K * 128 * 4;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tptr, 256);
  for (int idx = t; idx < N * K; idx += 128) {
    const int n = idx / K, k = idx % K;
    uint32_t hi, lo;
    split_tf32(W[n * K + k], hi, lo);
    *reinterpret_cast<uint32_t*>(Bhi + chunk_major_offset(n, k, N)) = hi;
    *reinterpret_cast<uint32_t*>(Blo + chunk_major_offset(n, k, N)) = lo;
  }
  if (mode >= 4) {          // 128B-swizzled operand: W[o][i] rows of 128 B (32 floats), 8-row atoms XOR-swizzled
    for (int idx = t; idx < N * K; idx += 128) {
      const int n = idx / K, k = idx % K;
      uint32_t hi, lo;
      split_tf32(W[n * K + k], hi, lo);
      const uint32_t off = (uint32_t)((k / 32) * (N * 128) + n * 128 + ((((k % 32) / 4) ^ (n % 8)) * 16) + (k % 4) * 4);
      *reinterpret_cast<uint32_t*>(Bhi + off) = hi;
      *reinterpret_cast<uint32_t*>(Blo + off) = lo;
    }
  }
  const bool ss = (mode == 1);
  if (ss) {
    for (int k = 0; k < K; ++k) {
      uint32_t hi, lo;
      split_tf32(A[t * K + k], hi, lo);
      *reinterpret_cast<uint32_t*>(Ahi + chunk_major_offset(t, k, 128)) = hi;
      *reinterpret_cast<uint32_t*>(Alo + chunk_major_offset(t, k, 128)) = lo;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tptr;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t colA_hi = 0, colA_lo = 64, colD = 128;
  if (!ss && mode < 3) {
    for (int c = 0; c < K; c += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) split_tf32(A[t * K + c + i], hi[i], lo[i]);
      tmem_st16(lane_addr + colA_hi + c, hi);
      tmem_st16(lane_addr + colA_lo + c, lo);
    }
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
  }
  if (mode == 4) {          // forward with the swizzled operand (K-major, SWIZZLE_128B)
    fence_proxy_async_smem();
    for (int c = 0; c < K; c += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) split_tf32(A[t * K + c + i], hi[i], lo[i]);
      tmem_st16(lane_addr + colA_hi + c, hi);
      tmem_st16(lane_addr + colA_lo + c, lo);
    }
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc_fence_after_sync();
      const uint32_t idesc = idesc_tf32(128, N);
      const uint64_t sw = (uint64_t)2 << 61;
      for (int j = 0; j < K / 8; ++j) {
        const uint32_t o = (j / 4) * (N * 128) + (j % 4) * 32;
        const uint64_t bh = smem_desc_kmajor(smem_u32(Bhi) + o, 16, 1024) | sw;
        const uint64_t bl = smem_desc_kmajor(smem_u32(Blo) + o, 16, 1024) | sw;
        mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bh, idesc, j > 0);
        mma_tf32_ts(tbase + colD, tbase + colA_lo + 8 * j, bh, idesc, 1);
        mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bl, idesc, 1);
      }
      mma_commit(mbar);
    }
    mbar_wait(mbar, 0);
    tc_fence_after_sync();
    for (int c = 0; c < N; c += 16) {
      float r[16];
      tmem_ld16(lane_addr + colD + c, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) D[t * N + c + i] = r[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
    return;
  }
  if (mode == 3 || mode == 5) {          // transposed use of the SAME smem operand: D2[128 x K] = G[128 x N] * W[N x K]
    for (int c = 0; c < N; c += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) split_tf32(A[t * N + c + i], hi[i], lo[i]);
      tmem_st16(lane_addr + colA_hi + c, hi);
      tmem_st16(lane_addr + colA_lo + c, lo);
    }
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
      tc_fence_after_sync();
      const uint32_t idesc = idesc_tf32(128, K) | (1u << 16);
      for (int j = 0; j < N / 8; ++j) {
        uint32_t lbo = 128, sbo = N * 16, start = j * 128;
        if (variant & 1) { lbo = N * 16; sbo = 128; }
        uint64_t extra = 0;
        if (variant & 2) extra = (uint64_t)1 << 52;
        if (mode == 5) {       // SWIZZLE_128B, MN-major: 8 rows (o) of 128 B per K-step; 32-wide MN groups LBO apart
          start = j * 1024; lbo = N * 128; sbo = 1024; extra = (uint64_t)2 << 61;
          if (variant & 1) { lbo = 1024; sbo = N * 128; }
        }
        const uint64_t bh = smem_desc_kmajor(smem_u32(Bhi) + start, lbo, sbo) | extra;
        const uint64_t bl = smem_desc_kmajor(smem_u32(Blo) + start, lbo, sbo) | extra;
        mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bh, idesc, j > 0);
        mma_tf32_ts(tbase + colD, tbase + colA_lo + 8 * j, bh, idesc, 1);
        mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bl, idesc, 1);
      }
      mma_commit(mbar);
    }
    mbar_wait(mbar, 0);
    tc_fence_after_sync();
    for (int c = 0; c < K; c += 16) {
      float r[16];
      tmem_ld16(lane_addr + colD + c, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) D[t * K + c + i] = r[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
    return;
  }
  if (t == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = idesc_tf32(128, N);
    for (int j = 0; j < K / 8; ++j) {
      const uint64_t bh = smem_desc_kmajor(smem_u32(Bhi) + 2 * j * N * 16, N * 16, 128);
      const uint64_t bl = smem_desc_kmajor(smem_u32(Blo) + 2 * j * N * 16, N * 16, 128);
      if (ss) {
        const uint64_t ah = smem_desc_kmajor(smem_u32(Ahi) + 2 * j * 128 * 16, 128 * 16, 128);
        const uint64_t al = smem_desc_kmajor(smem_u32(Alo) + 2 * j * 128 * 16, 128 * 16, 128);
        mma_tf32_ss(tbase + colD, ah, bh, idesc, j > 0);
        mma_tf32_ss(tbase + colD, al, bh, idesc, 1);
        mma_tf32_ss(tbase + colD, ah, bl, idesc, 1);
      } else {
        mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bh, idesc, j > 0);
        if (mode == 0) {
          mma_tf32_ts(tbase + colD, tbase + colA_lo + 8 * j, bh, idesc, 1);
          mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bl, idesc, 1);
        }
      }
    }
    mma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    float r[16];
    tmem_ld16(lane_addr + colD + c, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[t * N + c + i] = r[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

int main() {
  const int cases[3][2] = {{32, 64}, {64, 64}, {64, 32}};
  int fails = 0;
  for (int ci = 0; ci < 3; ++ci) {
    const int K = cases[ci][0], N = cases[ci][1];
    std::vector<float> A(128 * 64), W(N * K), D(128 * 64);
    srand(7 + ci);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    float *dA, *dW, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 1024 + 2 * K * N * 4 + 2 * K * 128 * 4;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mv = 0; mv < 8; ++mv) {
      const int modes[8] = {0, 1, 2, 3, 3, 4, 5, 5}, variants[8] = {0, 0, 0, 0, 1, 0, 0, 1};
      const int mode = modes[mv], variant = variants[mv];
      cudaMemset(dD, 0, D.size() * 4);
      probe_kernel<<<1, 128, smem>>>(dA, dW, dD, K, N, mode, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("K=%d N=%d mode=%d CUDA error %s\n", K, N, mode, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      if (mode == 3 || mode == 5) {
        for (int m = 0; m < 128; ++m)
          for (int k = 0; k < K; ++k) {
            double ref = 0;
            for (int n = 0; n < N; ++n) ref += (double)A[m * N + n] * (double)W[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - (double)D[m * K + k]));
            maxref = fmax(maxref, fabs(ref));
          }
      } else
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[n * K + k];
          maxerr = fmax(maxerr, fabs(ref - (double)D[m * N + n]));
          maxref = fmax(maxref, fabs(ref));
        }
      const char* names[6] = {"TS 3xTF32", "SS 3xTF32", "TS 1xTF32", "TS^T 3xTF32", "TS sw128", "TS^T sw128"};
      if (mode == 3 || mode == 5) {
        double r0 = 0, r1 = 0;
        for (int n = 0; n < N; ++n) { r0 += (double)A[0 * N + n] * W[n * K + 0]; r1 += (double)A[1 * N + n] * W[n * K + 5]; }
        printf("   D[0][0]=%g (ref %g)  D[1][5]=%g (ref %g)  D[127][K-1]=%g\n", D[0], r0, D[1 * K + 5], r1, D[127 * K + K - 1]);
      }
      const double rel = maxerr / maxref;
      const bool ok = (mode == 2) ? (rel < 5e-3) : (rel < 2e-6);
      printf("K=%2d N=%2d %-10s v%d max rel err %.3e  %s\n", K, N, names[mode], variant, rel, ok ? "OK" : "FAIL");
      fails += !ok;
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dD);
  }
  printf(fails ? "PROBE_FAIL\n" : "PROBE_OK\n");
  return fails ? 1 : 0;
}
