// fused_tc.cuh -- fused small-system kernels with the MLP contraction on the 5th-gen tensor cores.
//
// Tile = 128 frames = 128 threads = 128 TMEM lanes: thread t owns frame t of the tile end to end.
//   1. TMA bulk copy stages the tile's coordinates (one contiguous byte range of x) into smem;
//   2. thread-per-frame Kabsch (Jacobi on Horn's 4x4) + feature program -> feature row;
//   3. the row is split into TF32 hi/lo and written to TMEM as the A operand (tcgen05.st) -- activations
//      never touch shared memory; weights (hi/lo, chunk-major K-major) sit in smem for the whole kernel;
//   4. one thread issues tcgen05.mma (3 per K-step: hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM),
//      commits to an mbarrier; every thread reads its accumulator row back (tcgen05.ld), applies
//      bias + activation, splits, and stores the next layer's A operand;
//   5. the last (narrow) layer is a register dot product; y is written once.
// Two CTAs are resident per SM (256 TMEM columns each), so one CTA's tensor work overlaps the other's
// CUDA-core geometry/activation work.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "tc.cuh"
#include "fused_small.cuh"

namespace molann {

constexpr int TC_F = 128;          // frames per tile == threads per CTA == TMEM lanes
constexpr int TC_MAXW = 64;        // widest feature / hidden layer handled by this kernel
constexpr int TC_TMEM_COLS = 256;  // three 64-column regions (A_hi, A_lo, D) rounded up to a power of two

struct TcLayout {
  int xs_off, feat_off;              // coordinate tile [F][3n]; feature staging [Kp0][F]
  int bhi_off[MOLANN_MAX_LAYERS];    // per MMA layer: weights hi / lo, chunk-major [Kp/4][Np][4]
  int blo_off[MOLANN_MAX_LAYERS];
  int bias_off[MOLANN_MAX_LAYERS];   // padded biases
  int kp[MOLANN_MAX_LAYERS];         // padded K (multiple of 8)
  int np[MOLANN_MAX_LAYERS];         // padded N (multiple of 16)
  int wlast_off, blast_off;          // last layer: natural [k_out][K_last] fp32 + bias
  int aidx_off, ref_off, ent_off, mbar_off, tptr_off;
  int total_bytes;
};

// tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e)): 2 MUFU + 3 FMA-pipe ops, abs error ~2e-7
__device__ __forceinline__ float fast_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ float act_forward_fast(float v, int act) {
  switch (act) {
    case ACT_TANH: return fast_tanh(v);
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: {
      float e, r;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
      return r;
    }
    default: return v;
  }
}

// round-to-nearest TF32 split: |x - hi - lo| <= 2^-22 |x| once the MMA truncates lo to TF32
__device__ __forceinline__ void split_tf32_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// Split weights of MMA layer k into chunk-major hi/lo smem operands (zero padded to [np][kp]).
__device__ __forceinline__ void stage_tc_weights(const float* __restrict__ Wg, const float* __restrict__ bg, int K,
                                                 int N, int kp, int np, unsigned char* bhi, unsigned char* blo,
                                                 float* bias, int tid, int nthreads) {
  for (int idx = tid; idx < np * kp; idx += nthreads) {
    const int n = idx / kp, k = idx - n * kp;
    const float w = (n < N && k < K) ? Wg[(long long)n * K + k] : 0.f;
    uint32_t hi, lo;
    split_tf32_rn(w, hi, lo);
    lo = (lo + 0x1000u) & 0xffffe000u;
    const uint32_t off = chunk_major_offset(n, k, np);
    *reinterpret_cast<uint32_t*>(bhi + off) = hi;
    *reinterpret_cast<uint32_t*>(blo + off) = lo;
  }
  for (int n = tid; n < np; n += nthreads) bias[n] = (n < N) ? bg[n] : 0.f;
}

// One thread: D[128 x np] = A[128 x kp] * B^T with the 3xTF32 expansion, then commit to `mbar`.
__device__ __forceinline__ void issue_layer_mma(uint32_t tbase, uint32_t colA_hi, uint32_t colA_lo, uint32_t colD,
                                                const unsigned char* bhi, const unsigned char* blo, int kp, int np,
                                                void* mbar) {
  const uint32_t idesc = idesc_tf32(TC_F, np);
  const uint32_t bhi_a = smem_u32(bhi), blo_a = smem_u32(blo);
  const uint32_t step = 2u * (uint32_t)np * 16u;        // two 16-byte K-chunks per MMA (K = 8)
  const uint32_t lbo = (uint32_t)np * 16u;
  for (int j = 0; j < kp / 8; ++j) {
    const uint64_t bh = smem_desc_kmajor(bhi_a + j * step, lbo, 128);
    const uint64_t bl = smem_desc_kmajor(blo_a + j * step, lbo, 128);
    mma_tf32_ts(tbase + colD, tbase + colA_lo + 8 * j, bh, idesc, j > 0);      // small terms first
    mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bl, idesc, 1);
    mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bh, idesc, 1);
  }
  mma_commit(mbar);
}

template <int NT>
__device__ __forceinline__ void stage_tc_consts(const DevPlan& p, const TcLayout& lay, unsigned char* smem, int tid) {
  int* aidx = reinterpret_cast<int*>(smem + lay.aidx_off);
  float* ref = reinterpret_cast<float*>(smem + lay.ref_off);
  int* ent = reinterpret_cast<int*>(smem + lay.ent_off);
  for (int i = tid; i < p.n_align; i += NT) aidx[i] = p.align_idx[i];
  for (int i = tid; i < 3 * p.n_align; i += NT) ref[i] = p.ref_x[i];
  for (int i = tid; i < ENTRY_INTS * p.n_entries; i += NT) ent[i] = p.entries[i];
}

// =============================================================================================
// Forward
// =============================================================================================
__global__ void __launch_bounds__(TC_F)
fused_tc_forward_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ TcLayout lay,
                        const float* __restrict__ x, float* __restrict__ y, long long L, int use_tma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n3 = 3 * p.n_inp;
  const int nl = p.n_layers;          // layers 0 .. nl-2 run on tensor cores, layer nl-1 in registers
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  float* featbuf = reinterpret_cast<float*>(smem + lay.feat_off);
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  unsigned long long* mbar_x = reinterpret_cast<unsigned long long*>(smem + lay.mbar_off);
  unsigned long long* mbar_mma = mbar_x + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + lay.tptr_off);

  stage_tc_consts<TC_F>(p, lay, smem, tid);
  for (int k = 0; k < nl - 1; ++k)
    stage_tc_weights(p.W[k], p.b[k], p.dims[k], p.dims[k + 1], lay.kp[k], lay.np[k], smem + lay.bhi_off[k],
                     smem + lay.blo_off[k], reinterpret_cast<float*>(smem + lay.bias_off[k]), tid, TC_F);
  {
    const int K = p.dims[nl - 1], N = p.dims[nl];
    float* wl = reinterpret_cast<float*>(smem + lay.wlast_off);
    float* bl = reinterpret_cast<float*>(smem + lay.blast_off);
    for (int i = tid; i < N * TC_MAXW; i += TC_F) {
      const int o = i / TC_MAXW, j = i - o * TC_MAXW;
      wl[i] = (j < K) ? p.W[nl - 1][(long long)o * K + j] : 0.f;
    }
    for (int o = tid; o < N; o += TC_F) bl[o] = p.b[nl - 1][o];
  }
  for (int i = tid; i < lay.kp[0] * TC_F; i += TC_F) featbuf[i] = 0.f;      // padded feature rows stay zero
  if (tid == 0) {
    mbar_init(mbar_x, 1);
    mbar_init(mbar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tptr, TC_TMEM_COLS);
  fence_proxy_async_smem();           // weight operands are read by the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tptr;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t COL_AHI = 0, COL_ALO = 64, COL_D = 128;

  const long long ntiles = (L + TC_F - 1) / TC_F;
  const uint32_t tile_bytes = (uint32_t)TC_F * (uint32_t)n3 * 4u;
  uint32_t phase_x = 0, phase_m = 0;
  auto is_tma_tile = [&](long long t) { return use_tma && (t + 1) * (long long)TC_F <= L; };
  auto issue_x = [&](long long t) {
    if (tid == 0) {
      mbar_expect_tx(mbar_x, tile_bytes);
      bulk_g2s(xs, x + t * (long long)TC_F * n3, tile_bytes, mbar_x);
    }
  };
  long long tile = blockIdx.x;
  if (tile < ntiles && is_tma_tile(tile)) issue_x(tile);

  for (; tile < ntiles; tile += gridDim.x) {
    const long long f_base = tile * (long long)TC_F;
    const int nf = (int)((L - f_base) < (long long)TC_F ? (L - f_base) : (long long)TC_F);
    if (is_tma_tile(tile)) {
      mbar_wait(mbar_x, phase_x);
      phase_x ^= 1u;
    } else {
      const float* src = x + f_base * n3;
      for (int i = tid; i < nf * n3; i += TC_F) xs[i] = src[i];
      __syncthreads();
    }
    // ---- geometry: thread t <-> frame t ----
    {
      const int f = tid < nf ? tid : nf - 1;
      const float* xf = xs + f * n3;
      Rigid rg;
      const bool aligned = p.n_align > 0;
      if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      TileOut out{featbuf, tid, TC_F};
      for (int e = 0; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    // own feature row -> TF32 hi/lo -> TMEM (A operand of layer 0)
    __syncwarp();
    for (int c = 0; c < lay.kp[0]; c += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) split_tf32_rn(featbuf[(c + i) * TC_F + tid], hi[i], lo[i]);
      tmem_st8(lane_addr + COL_AHI + c, hi);
      tmem_st8(lane_addr + COL_ALO + c, lo);
    }
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    const long long next = tile + gridDim.x;
    if (next < ntiles && is_tma_tile(next)) issue_x(next);        // xs is free: overlap with the MLP
    if (tid == 0) {
      tc_fence_after_sync();
      issue_layer_mma(tbase, COL_AHI, COL_ALO, COL_D, smem + lay.bhi_off[0], smem + lay.blo_off[0], lay.kp[0],
                      lay.np[0], mbar_mma);
    }
    float h[TC_MAXW];
    for (int k = 0; k < nl - 1; ++k) {
      mbar_wait(mbar_mma, phase_m);
      phase_m ^= 1u;
      tc_fence_after_sync();
      const float* bias = reinterpret_cast<const float*>(smem + lay.bias_off[k]);
      const int np = lay.np[k];
      __syncwarp();
#pragma unroll
      for (int c = 0; c < TC_MAXW; c += 16)
        if (c < np) tmem_ld16(lane_addr + COL_D + c, *reinterpret_cast<float(*)[16]>(&h[c]));
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < TC_MAXW; c += 16) {
        if (c < np) {
#pragma unroll
          for (int i = 0; i < 16; ++i) h[c + i] = act_forward_fast(h[c + i] + bias[c + i], p.act);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) h[c + i] = 0.f;
        }
      }
      if (k < nl - 2) {          // feed the next tensor-core layer
#pragma unroll
        for (int c = 0; c < TC_MAXW; c += 16) {
          if (c < lay.kp[k + 1]) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_tf32_rn(h[c + i], hi[i], lo[i]);
            tmem_st16(lane_addr + COL_AHI + c, hi);
            tmem_st16(lane_addr + COL_ALO + c, lo);
          }
        }
        tmem_wait_st();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after_sync();
          issue_layer_mma(tbase, COL_AHI, COL_ALO, COL_D, smem + lay.bhi_off[k + 1], smem + lay.blo_off[k + 1],
                          lay.kp[k + 1], lay.np[k + 1], mbar_mma);
        }
      }
    }
    // ---- last layer in registers: y[o] = b[o] + sum_j h[j] W[o][j] ----
    {
      const int N = p.dims[nl];
      const float* wl = reinterpret_cast<const float*>(smem + lay.wlast_off);
      const float* bl = reinterpret_cast<const float*>(smem + lay.blast_off);
      float* yrow = y + (f_base + tid) * N;
      for (int o = 0; o < N; ++o) {
        float acc0 = bl[o], acc1 = 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(wl + o * TC_MAXW);
#pragma unroll
        for (int j = 0; j < TC_MAXW / 4; ++j) {
          const float4 w = w4[j];
          acc0 = fmaf(h[4 * j], w.x, acc0);
          acc1 = fmaf(h[4 * j + 1], w.y, acc1);
          acc0 = fmaf(h[4 * j + 2], w.z, acc0);
          acc1 = fmaf(h[4 * j + 3], w.w, acc1);
        }
        if (tid < nf) yrow[o] = acc0 + acc1;
      }
    }
    // the next tile's tcgen05.st / MMA reuse the TMEM regions: every thread's loads are complete
    // (tcgen05.wait::ld above) before it can pass the next __syncthreads.
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TC_TMEM_COLS);
}

}  // namespace molann
