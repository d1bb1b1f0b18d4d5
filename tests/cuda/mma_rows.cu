// mma_rows.cu -- DEVELOPMENT TOOL: does a tcgen05.mma.kind::tf32 result row depend (at rounding level) on WHERE
// the row sits in the 128-row tile?  All rows of A are made identical; every D row must then be bitwise equal.
// Checks the SS form (A in shared memory) and the TS form (A in TMEM) with the kernels' 3xTF32 sequence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rows mma_rows.cu && ./mma_rows
#include <cstdio>
#include <cstring>
#include "../../molann_b200/csrc/tc.cuh"
using namespace molann;

__device__ __forceinline__ void split_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// K = 32, N = 64
__global__ void rows_kernel(const float* arow, const float* W, float* dss, float* dts, int distinct) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* ahi = smem;                    // [8 chunks][128 rows][16 B]
  unsigned char* alo = smem + 16384;
  unsigned char* bhi = smem + 32768;            // [8 chunks][64 rows][16 B]
  unsigned char* blo = smem + 32768 + 8192;
  for (int k = 0; k < 32; ++k) {
    uint32_t hi, lo;
    split_rn(arow[k] * (distinct ? (1.0f + 0.37f * tid) : 1.0f), hi, lo);
    const int off = (k >> 2) * 2048 + tid * 16 + (k & 3) * 4;
    *reinterpret_cast<uint32_t*>(ahi + off) = hi;
    *reinterpret_cast<uint32_t*>(alo + off) = lo;
  }
  for (int idx = tid; idx < 64 * 32; idx += 128) {
    const int n = idx / 32, k = idx % 32;
    uint32_t hi, lo;
    split_rn(W[n * 32 + k], hi, lo);
    lo = (lo + 0x1000u) & 0xffffe000u;
    const uint32_t off = chunk_major_offset(n, k, 64);
    *reinterpret_cast<uint32_t*>(bhi + off) = hi;
    *reinterpret_cast<uint32_t*>(blo + off) = lo;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tptr, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tptr;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  // A in TMEM as well: hi cols [0,32), lo cols [32,64)
  for (int k = 0; k < 32; ++k) {
    uint32_t hi, lo;
    split_rn(arow[k] * (distinct ? (1.0f + 0.37f * tid) : 1.0f), hi, lo);
    tmem_st1(lane_addr + k, hi);
    tmem_st1(lane_addr + 32 + k, lo);
  }
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  const uint32_t idesc = idesc_tf32(128, 64);
  if (tid == 0) {
    tc_fence_after_sync();
    for (int j = 0; j < 4; ++j) {
      const uint64_t bh = smem_desc_kmajor(smem_u32(bhi) + j * 2048, 1024, 128);
      const uint64_t bl = smem_desc_kmajor(smem_u32(blo) + j * 2048, 1024, 128);
      const uint64_t ah = smem_desc_kmajor(smem_u32(ahi) + j * 4096, 2048, 128);
      const uint64_t al = smem_desc_kmajor(smem_u32(alo) + j * 4096, 2048, 128);
      mma_tf32_ss(tbase + 64, al, bh, idesc, j > 0);
      mma_tf32_ss(tbase + 64, ah, bl, idesc, 1);
      mma_tf32_ss(tbase + 64, ah, bh, idesc, 1);
      mma_tf32_ts(tbase + 128, tbase + 32 + 8 * j, bh, idesc, j > 0);
      mma_tf32_ts(tbase + 128, tbase + 8 * j, bl, idesc, 1);
      mma_tf32_ts(tbase + 128, tbase + 8 * j, bh, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  float r[16];
  for (int c = 0; c < 64; c += 16) {
    tmem_ld16(lane_addr + 64 + c, r);
    tmem_wait_ld();
    for (int i = 0; i < 16; ++i) dss[tid * 64 + c + i] = r[i];
    tmem_ld16(lane_addr + 128 + c, r);
    tmem_wait_ld();
    for (int i = 0; i < 16; ++i) dts[tid * 64 + c + i] = r[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

int main() {
  float ha[32], hw[64 * 32];
  unsigned s = 7u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  for (auto& v : ha) v = 8.0f * rnd();
  for (auto& v : hw) v = rnd();
  float *da, *dw, *dss, *dts;
  cudaMalloc(&da, sizeof(ha)); cudaMalloc(&dw, sizeof(hw)); cudaMalloc(&dss, 128 * 64 * 4); cudaMalloc(&dts, 128 * 64 * 4);
  cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice); cudaMemcpy(dw, hw, sizeof(hw), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  rows_kernel<<<1, 128, 64 * 1024>>>(da, dw, dss, dts, 1);
  cudaDeviceSynchronize();
  {
    static float a[128 * 64], b[128 * 64];
    cudaMemcpy(a, dss, sizeof(a), cudaMemcpyDeviceToHost); cudaMemcpy(b, dts, sizeof(b), cudaMemcpyDeviceToHost);
    int nd = 0; double maxrel = 0;
    for (int r = 0; r < 128; ++r) {
      bool diff = false;
      for (int c = 0; c < 64; ++c) if (std::memcmp(&a[r * 64 + c], &b[r * 64 + c], 4)) { diff = true; double e = (a[r*64+c]-b[r*64+c]) / (double)b[r*64+c]; if (e<0) e=-e; if (e>maxrel) maxrel=e; }
      nd += diff;
      if (diff && nd <= 8) printf("  distinct rows: SS != TS at row %d\n", r);
    }
    printf("distinct rows: SS vs TS differ bitwise in %d of 128 rows (max rel %.3g)\n", nd, maxrel);
  }
  rows_kernel<<<1, 128, 64 * 1024>>>(da, dw, dss, dts, 0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  static float hss[128 * 64], hts[128 * 64];
  cudaMemcpy(hss, dss, sizeof(hss), cudaMemcpyDeviceToHost); cudaMemcpy(hts, dts, sizeof(hts), cudaMemcpyDeviceToHost);
  for (int form = 0; form < 2; ++form) {
    const float* d = form ? hts : hss;
    int ndiff_rows = 0; float maxd = 0;
    for (int r = 1; r < 128; ++r) {
      bool diff = false;
      for (int c = 0; c < 64; ++c) {
        if (std::memcmp(&d[r * 64 + c], &d[c], 4) != 0) { diff = true; float e = d[r * 64 + c] - d[c]; if (e < 0) e = -e; if (e > maxd) maxd = e; }
      }
      ndiff_rows += diff;
      if (diff && ndiff_rows <= 6) printf("  form %s: row %d differs from row 0\n", form ? "TS" : "SS", r);
    }
    printf("%s form: %d of 127 rows differ bitwise from row 0 (max abs diff %.3g); D[0][0..3] = %.7g %.7g %.7g %.7g\n",
           form ? "TS" : "SS", ndiff_rows, maxd, d[0], d[1], d[2], d[3]);
  }
  double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)ha[k] * hw[k];
  printf("fp64 reference D[0][0] = %.9g ; SS-TS difference at [0][0] = %.3g\n", ref, hss[0] - hts[0]);
  return 0;
}
