// fused_ws.cuh -- warp-specialised fused forward kernel for small systems (the C2 class:
// n_inp small, feature / hidden widths <= 64, <= 8 outputs, 1 or 2 hidden layers).
//
// One persistent CTA per SM; tiles of 128 frames flow through a pipeline of ROLES that run concurrently on
// different tiles and hand over through mbarriers and TMEM:
//
//   producer (1 thread)   TMA bulk copy of the tile's contiguous byte range of x into a 3-deep smem ring
//   G  (NG x 4 warps)     thread = frame: pivoted moments, rotation (polynomial path, Jacobi fallback),
//                         feature program; features leave as TF32 hi/lo rows of the layer-1 A operand, a
//                         canonical K-major tile in SHARED memory (A1 in smem frees the TMEM columns that let
//                         A2 and both accumulators be double-buffered).
//                         The geometry is one long dependent chain per frame, so NG warpgroups work on NG
//                         different tiles at once (it needs no TMEM until its last step)
//   MMA-1 (1 thread)      tcgen05.mma 3xTF32  D1 = A1 * W1^T      (A from TMEM, W in smem for the whole kernel)
//   E1 (2 x 4 warps)      thread = frame: D1 -> bias + activation -> TF32 hi/lo -> A2 operand in TMEM
//                         (two warpgroups alternate tiles; each owns one D1 and one A2 buffer)
//   MMA-2 (1 thread)      D2 = A2 * W2^T
//   E2 (2 x 4 warps)      thread = frame: D2 -> bias + activation -> last (narrow) layer as a register dot
//                         product -> y
// 28 warps = 7 per scheduler; registers are re-balanced between the roles with setmaxnreg.
//
// Why: the single-role kernel (fused_tc.cuh) is bound by instruction issue and dependency latency with two
// warps per scheduler (profiles/r1_c: issue 43 %, stall "wait" 35 %), and every tile serialises on the two
// MMA round trips.  Here each scheduler holds a G, an E1 and an E2 warp with different pipe mixes (FMA chains /
// MUFU-heavy), the MMA latency is hidden behind the other roles' work on other tiles, and no CTA-wide barrier
// is left in the steady state.
//
// TMEM map (512 columns): A2 double-buffered (2 x 2*kp1) and the accumulators D1, D2 double-buffered, so that the
// MMA of tile i+1 runs while the epilogue still reads tile i (with single accumulators the epilogue and its
// MMA simply alternated, and with a single A2 the epilogue waited for the previous tile's MMA:
// tests/cuda/ws_trace.cu).  A1 (double-buffered) lives in shared memory.
// Weights and biases are pre-multiplied by the activation's exponent scale (tanh: 2 log2 e) while they are
// staged, so the epilogue is  e = ex2(acc + b');  h = 1 - 2 / (1 + e).
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "tc.cuh"
#include "fused_tc.cuh"

namespace molann {

constexpr int WS_F = 128;
constexpr int WS_NG = 2;                     // geometry warpgroups (tiles in flight in the G stage)
constexpr int WS_XBUF = WS_NG + 1;           // coordinate-tile ring
constexpr int WS_NE = 2;                     // warpgroups per epilogue role (alternate tiles, one D buffer each)
constexpr int WS_WARPS = 4 * WS_NG + 8 * WS_NE + 4;   // G + E1 + E2 + control warpgroup (producer, 2 MMA issuers, idle)
constexpr int WS_THREADS = WS_WARPS * 32;
constexpr int WS_W_E1 = 4 * WS_NG, WS_W_E2 = WS_W_E1 + 4 * WS_NE, WS_W_PROD = WS_W_E2 + 4 * WS_NE,
              WS_W_MMA1 = WS_W_PROD + 1, WS_W_MMA2 = WS_W_PROD + 2;
// register budget per role (setmaxnreg; the kernel is compiled for 65536 / WS_THREADS = 72 per thread):
// 8 G warps x 96 + 16 E warps x 64 + 4 control warps x 56 = 64512 <= 65536
constexpr int WS_REGS_G = 96, WS_REGS_E = 64, WS_REGS_CTRL = 56;

// Development aid (tests/cuda/ws_trace.cu): per-role, per-tile clock64() stamps of CTA 0.  Compiled out of the product.
#ifdef MOLANN_WS_TRACE
__device__ long long g_ws_trace[8 * 64 * 8];
__device__ long long g_ws_life[256 * 2];       // per CTA: start / end (globaltimer ns)
#define WS_EVT(role, i, ev)                                                                    \
  do {                                                                                         \
    if (blockIdx.x == 0 && (i) < 64 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) & 3) == 0) \
      g_ws_trace[((role) * 64 + (i)) * 8 + (ev)] = clock64();                                  \
  } while (0)
#define WS_EVT_L0(role, i, ev)                                                        \
  do {                                                                                \
    if (blockIdx.x == 0 && (i) < 64) g_ws_trace[((role) * 64 + (i)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define WS_EVT(role, i, ev) \
  do {                      \
  } while (0)
#define WS_EVT_L0(role, i, ev) \
  do {                         \
  } while (0)
#endif

struct WsLayout {
  TcLayout base;                             // weights / biases / plan constants (xs_off, mbar_off unused)
  int xs_off[WS_XBUF];
  int n_xbuf;                                // coordinate-tile ring depth actually used (2 when smem is short)
  int mbar_off, tptr_off;
  int ref4_off;                              // reference rows padded to float4
  int aoff_off;                              // 3 * align_idx (element offsets into a frame)
  int n_a2buf;                               // 1 or 2 A2 buffers
  int n_dbuf;                                // 1 or 2 buffers per accumulator
  int a1s_off[2];                            // layer-1 A operand buffers in SHARED memory (hi block, lo block)
  int ones_off;                              // constant A tile [128 x 8]: column 0 = 1 (bias through the MMA)
  int bbh_off[2], bbl_off[2];                // per MMA layer: bias as a [np x 8] B operand (column 0), hi / lo
  int col_a2[2], col_d1[2], col_d2[2];       // TMEM column bases
  int tmem_cols;
  int total_bytes;
};

#ifndef MOLANN_SHARED_RCP
#define MOLANN_SHARED_RCP 0   // paired activation with one reciprocal (A/B switch)
#endif
#ifndef MOLANN_WAIT_NS
#define MOLANN_WAIT_NS 64     // back-off between two polls of an mbarrier (A/B switch of tests/cuda/ws_trace.cu)
#endif
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of burning issue slots that the working roles on the same scheduler need
__device__ __forceinline__ void mbar_wait_hint(void* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WS_DONE_%=;\n\t"
      "WS_WAIT_%=:\n\t"
      "nanosleep.u32 %3;\n\t"             // retries were a third of all issued instructions (profiles/r1_f)
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@!p bra WS_WAIT_%=;\n\t"
      "WS_DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u), "n"(MOLANN_WAIT_NS)
      : "memory");
}

__host__ __device__ inline float ws_scale_for_act(int act) {
  return act == ACT_TANH ? 2.8853900817779268f : (act == ACT_SIGMOID ? -1.4426950408889634f : 1.0f);
}

// activation on a pre-scaled pre-activation
template <int ACT>
__device__ __forceinline__ float ws_act(float zs) {
  if (ACT == ACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zs));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return fmaf(-2.0f, r, 1.0f);
  } else if (ACT == ACT_SIGMOID) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zs));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
  } else if (ACT == ACT_RELU) {
    return fmaxf(zs, 0.f);
  }
  return zs;
}

// the same activation on two pre-scaled pre-activations: the adds and the final 1 - 2 r as packed f32x2 operations
template <int ACT>
__device__ __forceinline__ void ws_act_x2(float z0, float z1, float& h0, float& h1) {
  if (ACT == ACT_TANH || ACT == ACT_SIGMOID) {
    float e0, e1, r0, r1;
#if MOLANN_SHARED_RCP
    // ONE reciprocal for the pair: 1 / s0 = s1 / (s0 s1) -- three MUFU operations per pair instead of four (the XU pipe
    // is 58 % busy in the C2 forward and the geometry role's rsqrt / div queue behind the epilogues' on it).  The
    // clamp keeps s0 s1 finite (2^124); tanh / sigmoid are saturated to fp32 rounding long before (|x| > 21).
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(z0, 62.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(z1, 62.0f)));
    float s0, s1, rp;
    f2_unpack(f2_add(f2_pack(e0, e1), f2_pack(1.0f, 1.0f)), s0, s1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rp) : "f"(s0 * s1));
    f2_unpack(f2_mul(f2_pack(rp, rp), f2_pack(s1, s0)), r0, r1);
#else
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
    float s0, s1;
    f2_unpack(f2_add(f2_pack(e0, e1), f2_pack(1.0f, 1.0f)), s0, s1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(s1));
#endif
    if (ACT == ACT_TANH) {
      f2_unpack(f2_fma(f2_pack(-2.0f, -2.0f), f2_pack(r0, r1), f2_pack(1.0f, 1.0f)), h0, h1);
    } else {
      h0 = r0;
      h1 = r1;
    }
  } else {
    h0 = ws_act<ACT>(z0);
    h1 = ws_act<ACT>(z1);
  }
}

// weights of MMA layer k, multiplied by `scale`, split into chunk-major hi/lo operands
__device__ __forceinline__ void ws_stage_weights(const float* __restrict__ Wg, const float* __restrict__ bg, int K, int N,
                                                 int kp, int np, float scale, unsigned char* bhi, unsigned char* blo,
                                                 float* bias, int tid, int nthreads) {
  // 4 elements per thread and step with the (L2-latency) loads issued together: this prologue is part of the
  // fixed cost of a launch (tests/cuda/ws_trace.cu: ~17 us before, of a 138 us bench step)
  for (int base = tid * 4; base < np * kp; base += nthreads * 4) {
    float w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = base + q;                 // kp is a multiple of 16: the 4 elements share a row and a K-chunk
      const int n = idx / kp, k = idx - n * kp;
      w[q] = (n < N && k < K) ? __ldg(Wg + (long long)n * K + k) : 0.f;
    }
    const int n = base / kp, k = base - n * kp;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      split_tf32_rn(scale * w[q], hi[q], lo[q]);
      lo[q] = (lo[q] + 0x1000u) & 0xffffe000u;
    }
    const uint32_t off = chunk_major_offset(n, k, np);
    *reinterpret_cast<uint4*>(bhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(blo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (bias != nullptr)
    for (int n = tid; n < np; n += nthreads) bias[n] = (n < N) ? scale * bg[n] : 0.f;
}

// Whole-layer issue, descriptors formed on the fly (the control warps run on a 40-register budget).
// TS form: A (hi / lo column blocks) in TMEM.  SS form: A in shared memory, canonical K-major tile
// (hi block at `a_smem`, lo block kp * 512 bytes further; K-step = two 2048-byte K-chunks).
// The bias enters through the tensor core as well: D = ONES[128 x 8] * BIAS[np x 8]^T (hi, then lo) starts the
// accumulation, which takes a shared-memory load and an add per activation out of both epilogues for two extra
// MMAs per layer.
__device__ __forceinline__ void ws_issue_bias(uint32_t leader, uint32_t d, const unsigned char* ones,
                                              const unsigned char* bbh, const unsigned char* bbl, int np, uint32_t idesc) {
  const uint64_t a1 = smem_desc_kmajor(smem_u32(ones), WS_F * 16u, 128);
  const uint64_t bh = smem_desc_kmajor(smem_u32(bbh), (uint32_t)np * 16u, 128);
  const uint64_t bl = smem_desc_kmajor(smem_u32(bbl), (uint32_t)np * 16u, 128);
  if (leader) {
    mma_tf32_ss(d, a1, bl, idesc, 0);
    mma_tf32_ss(d, a1, bh, idesc, 1);
  }
}

__device__ __forceinline__ void ws_issue_layer_ts(uint32_t leader, uint32_t a_hi, uint32_t a_lo, uint32_t d,
                                                  const unsigned char* bhi, const unsigned char* blo, int kp, int np) {
  const uint32_t idesc = idesc_tf32(WS_F, np);
  const uint32_t bhi_a = smem_u32(bhi), blo_a = smem_u32(blo);
  const uint32_t step = 2u * (uint32_t)np * 16u, lbo = (uint32_t)np * 16u;
#pragma unroll 2
  for (int j = 0; j < kp / 8; ++j) {
    const uint64_t bh = smem_desc_kmajor(bhi_a + j * step, lbo, 128);
    const uint64_t bl = smem_desc_kmajor(blo_a + j * step, lbo, 128);
    if (leader) {
      mma_tf32_ts(d, a_lo + 8 * j, bh, idesc, 1);          // small terms first (the bias MMAs opened the sum)
      mma_tf32_ts(d, a_hi + 8 * j, bl, idesc, 1);
      mma_tf32_ts(d, a_hi + 8 * j, bh, idesc, 1);
    }
  }
}
__device__ __forceinline__ void ws_issue_layer_ss(uint32_t leader, const unsigned char* a_smem, uint32_t d,
                                                  const unsigned char* bhi, const unsigned char* blo, int kp, int np) {
  const uint32_t idesc = idesc_tf32(WS_F, np);
  const uint32_t bhi_a = smem_u32(bhi), blo_a = smem_u32(blo);
  const uint32_t ahi_a = smem_u32(a_smem), alo_a = ahi_a + (uint32_t)kp * (WS_F * 4u);
  const uint32_t step = 2u * (uint32_t)np * 16u, lbo = (uint32_t)np * 16u;
#pragma unroll 2
  for (int j = 0; j < kp / 8; ++j) {
    const uint64_t bh = smem_desc_kmajor(bhi_a + j * step, lbo, 128);
    const uint64_t bl = smem_desc_kmajor(blo_a + j * step, lbo, 128);
    const uint64_t ah = smem_desc_kmajor(ahi_a + j * (2u * WS_F * 16u), WS_F * 16u, 128);
    const uint64_t al = smem_desc_kmajor(alo_a + j * (2u * WS_F * 16u), WS_F * 16u, 128);
    if (leader) {
      mma_tf32_ss(d, al, bh, idesc, 1);                    // small terms first (the bias MMAs opened the sum)
      mma_tf32_ss(d, ah, bl, idesc, 1);
      mma_tf32_ss(d, ah, bh, idesc, 1);
    }
  }
}

// ---- E1: accumulator -> activation -> next layer's A operand, 16 columns at a time -------------------
// Rolled loops: each role streams its own code and the roles of a scheduler share the instruction caches
// (profiles/r1_d: "no instruction" was the top stall with unrolled epilogues).
template <int ACT>
__device__ __forceinline__ void ws_hidden_epilogue(uint32_t lane_d, uint32_t lane_ahi, uint32_t lane_alo,
                                                   const float* __restrict__ bias, int np,
                                                   unsigned long long* bar_d_full, unsigned long long* bar_d_free,
                                                   uint32_t par_d, void* bar_a_empty, uint32_t par_a_empty,
                                                   void* bar_a_full, int it) {
#pragma unroll 1
  for (int c0 = 0; c0 < np; c0 += 16) {
    if (c0 == 0) {
      mbar_wait_hint(bar_d_full, par_d);
      tc_fence_after_sync();
      WS_EVT(2, it, 0);
    }
    float z[16];
    tmem_ld16(lane_d + c0, z);
    tmem_wait_ld();
    if (c0 + 16 >= np) {                       // the accumulator has been read: its buffer may be overwritten
      tc_fence_before_sync();
      mbar_arrive(bar_d_free);
    }
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int c = 0; c < 16; c += 2) {                                                 // bias is already in z
      float h0, h1;
      ws_act_x2<ACT>(z[c], z[c + 1], h0, h1);
      split_tf32_rn_x2(h0, h1, hi[c], hi[c + 1], lo[c], lo[c + 1]);
    }
    if (c0 == 0) {                             // the MMA that last read this A2 buffer is complete
      mbar_wait_hint(bar_a_empty, par_a_empty);
      tc_fence_after_sync();
      WS_EVT(2, it, 1);
    }
    tmem_st16(lane_ahi + c0, hi);
    tmem_st16(lane_alo + c0, lo);
  }
  tmem_wait_st();
  tc_fence_before_sync();
  mbar_arrive(bar_a_full);
  WS_EVT(2, it, 3);
}

// ---- E2: accumulator -> activation -> last layer (register dot products) -> y ---------------------
template <int ACT>
__device__ __forceinline__ void ws_final_epilogue(uint32_t lane_d, const float* __restrict__ bias, int np,
                                                  unsigned long long* bar_d_full, unsigned long long* bar_d_free,
                                                  uint32_t par_d, const float* __restrict__ wl,
                                                  const float* __restrict__ bl, int kout, float* __restrict__ yrow,
                                                  bool valid, int it) {
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = (o < kout) ? bl[o] : 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < np; c0 += 16) {
    if (c0 == 0) {
      mbar_wait_hint(bar_d_full, par_d);
      tc_fence_after_sync();
      WS_EVT(3, it, 0);
    }
    float z[16];
    tmem_ld16(lane_d + c0, z);
    tmem_wait_ld();
    if (c0 + 16 >= np) {
      tc_fence_before_sync();
      mbar_arrive(bar_d_free);
    }
#pragma unroll
    for (int c = 0; c < 16; c += 2) ws_act_x2<ACT>(z[c], z[c + 1], z[c], z[c + 1]);    // bias is already in z
    if (kout == 2) {                            // the common case: two collective variables
      // packed FMAs: lane x of a pair accumulates the even columns, lane y the odd ones (same order as before)
      const float4* w0 = reinterpret_cast<const float4*>(wl + c0);
      const float4* w1 = reinterpret_cast<const float4*>(wl + TC_MAXW + c0);
      unsigned long long p0 = f2_pack(0.f, 0.f), p1 = f2_pack(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 u = w0[j], v = w1[j];
        const unsigned long long za = f2_pack(z[4 * j], z[4 * j + 1]), zb = f2_pack(z[4 * j + 2], z[4 * j + 3]);
        p0 = f2_fma(za, f2_pack(u.x, u.y), p0);
        p0 = f2_fma(zb, f2_pack(u.z, u.w), p0);
        p1 = f2_fma(za, f2_pack(v.x, v.y), p1);
        p1 = f2_fma(zb, f2_pack(v.z, v.w), p1);
      }
      float a0, d0, a1, d1;
      f2_unpack(p0, a0, d0);
      f2_unpack(p1, a1, d1);
      acc[0] += a0 + d0;
      acc[1] += a1 + d1;
    } else {
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        if (o < kout) {
          const float4* w4 = reinterpret_cast<const float4*>(wl + o * TC_MAXW + c0);
          float a0 = 0.f, d0 = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 w = w4[j];
            a0 = fmaf(z[4 * j], w.x, a0); d0 = fmaf(z[4 * j + 1], w.y, d0);
            a0 = fmaf(z[4 * j + 2], w.z, a0); d0 = fmaf(z[4 * j + 3], w.w, d0);
          }
          acc[o] += a0 + d0;
        }
      }
    }
  }
  WS_EVT(3, it, 2);
  if (valid) {
    if (kout == 2) {
      *reinterpret_cast<float2*>(yrow) = make_float2(acc[0], acc[1]);
    } else {
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o < kout) yrow[o] = acc[o];
    }
  }
}

// ---- G: features -> layer-1 A operand in SHARED memory ---------------------------------------------
// The operand is the canonical K-major, no-swizzle tile: element (row r, column k) of a block lives at
// ((k / 4) * 128 + r) * 16 + (k % 4) * 4 bytes (8 x 16 B core matrices; LBO = 2048 B between K-chunks,
// SBO = 128 B between 8-row groups).  Thread r owns row r, so a warp's 16-byte stores to one K-chunk cover
// 512 contiguous bytes (conflict-free).  Keeping A1 out of TMEM is what lets A2 and both accumulators be
// double-buffered in the 512 TMEM columns.
struct SmemFeatOut {
  unsigned char* hi_row;       // block base + r * 16
  unsigned char* lo_row;
  __device__ __forceinline__ void operator()(int col, float v) {
    uint32_t hi, lo;
    split_tf32_rn(v, hi, lo);
    const int off = (col >> 2) * (WS_F * 16) + (col & 3) * 4;
    *reinterpret_cast<uint32_t*>(hi_row + off) = hi;
    *reinterpret_cast<uint32_t*>(lo_row + off) = lo;
  }
};

// z = (x - c) R = x R - t with t = c R, for NA position entries starting at entry e0 (3 NA columns from 3 e0)
template <int NA>
__device__ __forceinline__ void ws_position_group(const float* __restrict__ xf, const int* __restrict__ ent, int e0,
                                                  const float (&R)[9], float t0, float t1, float t2,
                                                  SmemFeatOut& out) {
  uint32_t hi[3 * NA], lo[3 * NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const float* p = xf + 3 * ent[ENTRY_INTS * (e0 + i) + 1];
    const float px = p[0], py = p[1], pz = p[2];
    const float zx = fmaf(px, R[0], fmaf(py, R[3], fmaf(pz, R[6], -t0)));
    const float zy = fmaf(px, R[1], fmaf(py, R[4], fmaf(pz, R[7], -t1)));
    const float zz = fmaf(px, R[2], fmaf(py, R[5], fmaf(pz, R[8], -t2)));
    split_tf32_rn(zx, hi[3 * i], lo[3 * i]);
    split_tf32_rn(zy, hi[3 * i + 1], lo[3 * i + 1]);
    split_tf32_rn(zz, hi[3 * i + 2], lo[3 * i + 2]);
  }
  // 3 NA columns starting at a multiple of 12 (chunk aligned); groups of fewer than 4 atoms are zero-padded to
  // whole 16-byte K-chunks (the zeros are operand padding, or are overwritten by the entries that follow) so that
  // every store is a conflict-free STS.128
  constexpr int NC = (3 * NA + 3) / 4;
  const int off = (3 * e0 >> 2) * (WS_F * 16);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    uint32_t h4[4], l4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h4[q] = (4 * c + q < 3 * NA) ? hi[4 * c + q] : 0u;
      l4[q] = (4 * c + q < 3 * NA) ? lo[4 * c + q] : 0u;
    }
    *reinterpret_cast<uint4*>(out.hi_row + off + c * (WS_F * 16)) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    *reinterpret_cast<uint4*>(out.lo_row + off + c * (WS_F * 16)) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
  }
}

// returns the first column not yet written
__device__ __forceinline__ int ws_position_features(const float* __restrict__ xf, const int* __restrict__ ent,
                                                    int n_lead, const Rigid& rg, SmemFeatOut& out) {
  const float t0 = fmaf(rg.c[0], rg.R[0], fmaf(rg.c[1], rg.R[3], rg.c[2] * rg.R[6]));
  const float t1 = fmaf(rg.c[0], rg.R[1], fmaf(rg.c[1], rg.R[4], rg.c[2] * rg.R[7]));
  const float t2 = fmaf(rg.c[0], rg.R[2], fmaf(rg.c[1], rg.R[5], rg.c[2] * rg.R[8]));
  int e0 = 0;
#pragma unroll 1
  for (; e0 + 4 <= n_lead; e0 += 4) ws_position_group<4>(xf, ent, e0, rg.R, t0, t1, t2, out);
  const int rem = n_lead - e0;
  if (rem == 3) ws_position_group<3>(xf, ent, e0, rg.R, t0, t1, t2, out);
  else if (rem == 2) ws_position_group<2>(xf, ent, e0, rg.R, t0, t1, t2, out);
  else if (rem == 1) ws_position_group<1>(xf, ent, e0, rg.R, t0, t1, t2, out);
  return (3 * n_lead + 3) & ~3;
}

// =============================================================================================
template <int ACT>
__global__ void __launch_bounds__(WS_THREADS, 1)
fused_ws_forward_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ WsLayout wl,
                        const float* __restrict__ x, float* __restrict__ y, long long L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const TcLayout& lay = wl.base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef MOLANN_WS_TRACE
  if (blockIdx.x == 0 && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g_ws_trace[7 * 64 * 8 + 0] = clock64();
    g_ws_trace[7 * 64 * 8 + 1] = (long long)gt;
  }
  if (tid == 0 && blockIdx.x < 256) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g_ws_life[blockIdx.x * 2] = (long long)gt;
  }
#endif
  const int n3 = 3 * p.n_inp;
  const int nl = p.n_layers;
  const int nh = nl - 1;                      // 1 or 2 tensor-core layers
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + wl.mbar_off);
  unsigned long long* x_full = bars;                   // [WS_XBUF]
  unsigned long long* x_empty = bars + WS_XBUF;        // [WS_XBUF]
  unsigned long long* a1_full = bars + 2 * WS_XBUF;    // [2]
  unsigned long long* a1_empty = a1_full + 2;          // [2]
  unsigned long long* a2_full = a1_empty + 2;          // [2]
  unsigned long long* a2_empty = a2_full + 2;          // [2]
  unsigned long long* d1_full = a2_empty + 2;          // [2]: accumulator buffers
  unsigned long long* d1_free = d1_full + 2;           // [2]
  unsigned long long* d2_full = d1_free + 2;           // [2]
  unsigned long long* d2_free = d2_full + 2;           // [2]
  unsigned long long* w_ready = d2_free + 2;           // weights / biases staged
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + wl.tptr_off);

  // ---- one-time staging.  What the geometry needs (a few hundred bytes) is staged by everyone before the first
  // barrier; the weights are staged by the 16 epilogue warps AFTER it, while the producer and the geometry warps
  // already work on the first tiles, and are published through the `w_ready` mbarrier. ----
  {
    int* aoff = reinterpret_cast<int*>(smem + wl.aoff_off);
    float4* ref4 = reinterpret_cast<float4*>(smem + wl.ref4_off);
    int* ent = reinterpret_cast<int*>(smem + lay.ent_off);
    for (int i = tid; i < p.n_align; i += WS_THREADS) {
      aoff[i] = 3 * p.align_idx[i];
      ref4[i] = make_float4(p.ref_x[3 * i], p.ref_x[3 * i + 1], p.ref_x[3 * i + 2], 0.f);
    }
    for (int i = tid; i < ENTRY_INTS * p.n_entries; i += WS_THREADS) ent[i] = p.entries[i];
  }
  if (tid == 0) {
    for (int b = 0; b < WS_XBUF; ++b) {
      mbar_init(&x_full[b], 1);
      mbar_init(&x_empty[b], WS_F);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a1_full[b], WS_F);
      mbar_init(&a1_empty[b], 1);
      mbar_init(&a2_full[b], WS_F);
      mbar_init(&a2_empty[b], 1);
    }
    mbar_init(w_ready, 8 * WS_NE * 32);
    for (int h = 0; h < 2; ++h) {
      mbar_init(&d1_full[h], 1);
      mbar_init(&d1_free[h], WS_F);
      mbar_init(&d2_full[h], 1);
      mbar_init(&d2_free[h], WS_F);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tptr, 512u);      // all of it (one CTA per SM): the base address is then 0
  fence_proxy_async_smem();                   // weight operands are read by the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tptr != 0u) __trap();                  // a full-TMEM allocation starts at column 0, lane 0
  // a literal base keeps every tcgen05 address warp-uniform for ptxas (uniform registers, no R2UR per MMA)
  constexpr uint32_t tbase = 0u;
  const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);

  const long long ntiles = (L + WS_F - 1) / WS_F;
  const uint32_t tile_bytes = (uint32_t)WS_F * (uint32_t)n3 * 4u;      // a multiple of 16
  const long long first = blockIdx.x, stride = gridDim.x;
  // x only has to be 4-byte aligned: the bulk copy starts at the 16-byte boundary below the tile (every tile has
  // the same misalignment because tile_bytes % 16 == 0) and the readers skip the first `xoff` bytes.  One kernel
  // for every alignment keeps a frame's result independent of how the batch was sliced.
  const uint32_t xoff = (uint32_t)(reinterpret_cast<uintptr_t>(x) & 15u);
  // with one hidden layer the only accumulator is "D2" and E1 / MMA-2 have nothing to do
  const int ndb = wl.n_dbuf;

  // (each setmaxnreg sits at the top of the code it governs: ptxas only applies the budget to a region the
  // instruction clearly dominates -- placed in one if-chain ahead of the role dispatch it was ignored and the
  // roles were allocated for the 72-register launch bound)
  if (warp >= WS_W_PROD) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REGS_CTRL));

  if (warp >= WS_W_E1 && warp < WS_W_PROD) {            // epilogue warps: stage weights, biases, last layer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REGS_E));
    const int st = tid - WS_W_E1 * 32, snt = 8 * WS_NE * 32;
    const float scale = ws_scale_for_act(ACT);
    for (int k = 0; k < nh; ++k)
      ws_stage_weights(p.W[k], p.b[k], p.dims[k], p.dims[k + 1], lay.kp[k], lay.np[k], scale, smem + lay.bhi_off[k],
                       smem + lay.blo_off[k], nullptr, st, snt);
    for (int i = st; i < 2 * WS_F; i += snt)         // ONES: chunk 0 of row r = (1, 0, 0, 0), chunk 1 = 0
      reinterpret_cast<float4*>(smem + wl.ones_off)[i] = make_float4(i < WS_F ? 1.f : 0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < nh; ++k) {                   // bias of layer k as a [np x 8] K-major operand, column 0
      const int np = lay.np[k], N = p.dims[k + 1];
      for (int i = st; i < 2 * np; i += snt) {
        uint32_t hi = 0u, lo = 0u;
        if (i < N) {
          split_tf32_rn(scale * p.b[k][i], hi, lo);
          lo = (lo + 0x1000u) & 0xffffe000u;
        }
        reinterpret_cast<uint4*>(smem + wl.bbh_off[k])[i] = make_uint4(i < np ? hi : 0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(smem + wl.bbl_off[k])[i] = make_uint4(i < np ? lo : 0u, 0u, 0u, 0u);
      }
    }
    const int K = p.dims[nl - 1], N = p.dims[nl];
    float* wlast = reinterpret_cast<float*>(smem + lay.wlast_off);
    float* blast = reinterpret_cast<float*>(smem + lay.blast_off);
    for (int i = st; i < N * TC_MAXW; i += snt) {
      const int o = i / TC_MAXW, j = i - o * TC_MAXW;
      wlast[i] = (j < K) ? p.W[nl - 1][(long long)o * K + j] : 0.f;
    }
    for (int o = st; o < N; o += snt) blast[o] = p.b[nl - 1][o];
    fence_proxy_async_smem();                  // weight operands are read by the tensor core (async proxy)
    mbar_arrive(w_ready);
    mbar_wait_hint(w_ready, 0u);
  }

  if (warp == WS_W_PROD) {
    // ================= producer =================
    int i = 0;
    for (long long tile = first; tile < ntiles; tile += stride, ++i) {
      const int b = i % wl.n_xbuf;
      const uint32_t par = (uint32_t)((i / wl.n_xbuf) & 1);
      float* dst = reinterpret_cast<float*>(smem + wl.xs_off[b]);
      const long long f_base = tile * (long long)WS_F;
      // a misaligned copy reads up to 16 bytes past the tile: fine inside the batch, not on its last tile
      const bool full = f_base + WS_F <= L && (xoff == 0u || f_base + WS_F < L);
      if (lane == 0) {
        mbar_wait_hint(&x_empty[b], par ^ 1u);
        WS_EVT_L0(6, i, 0);
      }
      __syncwarp();
      if (full) {
        if (lane == 0) {
          const uint32_t bytes = (tile_bytes + xoff + 15u) & ~15u;
          mbar_expect_tx(&x_full[b], bytes);
          bulk_g2s(dst, reinterpret_cast<const unsigned char*>(x + f_base * n3) - xoff, bytes, &x_full[b]);
        }
      } else {                                 // last tile (ragged, or misaligned): plain coalesced copy by the warp
        const long long rest = L - f_base;
        const int nf = (int)(rest < (long long)WS_F ? rest : (long long)WS_F);
        const float* src = x + f_base * n3;
        float* dsto = dst + (xoff >> 2);
        for (int j = lane; j < nf * n3; j += 32) dsto[j] = src[j];
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_full[b]);
      }
    }
  } else if (warp == WS_W_MMA1) {
    // ================= MMA issuer, layer 1 (whole warp runs the loop, one elected lane issues) =================
    {
      const uint32_t leader = elect_one();
      mbar_wait_hint(w_ready, 0u);
      int i = 0;
      for (long long tile = first; tile < ntiles; tile += stride, ++i) {
        const int ab = i & 1;
        const int db = i % ndb;
        const uint32_t dpar = (uint32_t)((i / ndb) & 1);
        unsigned long long* dfree = (nh == 2 ? d1_free : d2_free) + db;
        unsigned long long* dfull = (nh == 2 ? d1_full : d2_full) + db;
        const uint32_t colD = (uint32_t)(nh == 2 ? wl.col_d1[db] : wl.col_d2[db]);
        mbar_wait_hint(&a1_full[ab], (uint32_t)((i >> 1) & 1));
        if (leader) WS_EVT_L0(4, i, 0);
        mbar_wait_hint(dfree, dpar ^ 1u);
        tc_fence_after_sync();
        if (leader) WS_EVT_L0(4, i, 1);
        ws_issue_bias(leader, tbase + colD, smem + wl.ones_off, smem + wl.bbh_off[0], smem + wl.bbl_off[0], lay.np[0],
                      idesc_tf32(WS_F, lay.np[0]));
        ws_issue_layer_ss(leader, smem + wl.a1s_off[ab], tbase + colD, smem + lay.bhi_off[0], smem + lay.blo_off[0],
                          lay.kp[0], lay.np[0]);
        if (leader) mma_commit(dfull);
        if (leader) WS_EVT_L0(4, i, 2);
        if (leader) mma_commit(&a1_empty[ab]);
        __syncwarp();
      }
    }
  } else if (warp == WS_W_MMA2) {
    // ================= MMA issuer, layer 2 =================
    if (nh == 2) {
      const uint32_t leader = elect_one();
      mbar_wait_hint(w_ready, 0u);
      int i = 0;
      for (long long tile = first; tile < ntiles; tile += stride, ++i) {
        const int ab = i % wl.n_a2buf;
        const uint32_t colA = (uint32_t)wl.col_a2[ab];
        mbar_wait_hint(&a2_full[ab], (uint32_t)((i / wl.n_a2buf) & 1));
        if (leader) WS_EVT_L0(5, i, 0);
        const int db = i % ndb;
        mbar_wait_hint(&d2_free[db], (uint32_t)(((i / ndb) & 1) ^ 1));
        tc_fence_after_sync();
        if (leader) WS_EVT_L0(5, i, 1);
        ws_issue_bias(leader, tbase + (uint32_t)wl.col_d2[db], smem + wl.ones_off, smem + wl.bbh_off[1],
                      smem + wl.bbl_off[1], lay.np[1], idesc_tf32(WS_F, lay.np[1]));
        ws_issue_layer_ts(leader, tbase + colA, tbase + colA + lay.kp[1], tbase + (uint32_t)wl.col_d2[db],
                          smem + lay.bhi_off[1], smem + lay.blo_off[1], lay.kp[1], lay.np[1]);
        if (leader) mma_commit(&d2_full[db]);
        if (leader) WS_EVT_L0(5, i, 2);
        if (leader) mma_commit(&a2_empty[ab]);
        __syncwarp();
      }
    }
  } else if (warp < WS_W_E1) {
    // ================= G: geometry + features; warpgroup g takes local tiles g, g + NG, ... =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REGS_G));
    const int g = warp >> 2;
    const int ft = tid & (WS_F - 1);            // frame within the tile
    const int* aoff = reinterpret_cast<const int*>(smem + wl.aoff_off);
    const float4* ref4 = reinterpret_cast<const float4*>(smem + wl.ref4_off);
    const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
    const bool aligned = p.n_align > 0;
    const int kp0 = lay.kp[0];
    int n_lead = 0;                              // leading position entries take the unrolled path
    while (n_lead < p.n_entries && ent[ENTRY_INTS * n_lead] == FEAT_POSITION) ++n_lead;
    const bool mixed = n_lead < p.n_entries;
    int i = g;
    for (long long tile = first + (long long)g * stride; tile < ntiles; tile += WS_NG * stride, i += WS_NG) {
      const int b = i % wl.n_xbuf;
      const long long f_base = tile * (long long)WS_F;
      const int nf = (int)((L - f_base) < (long long)WS_F ? (L - f_base) : (long long)WS_F);
      const int f = ft < nf ? ft : nf - 1;
      const float* xf = reinterpret_cast<const float*>(smem + wl.xs_off[b] + xoff) + f * n3;
      mbar_wait_hint(&x_full[b], (uint32_t)((i / wl.n_xbuf) & 1));
      WS_EVT(g, i, 0);
      Rigid rg;
      if (aligned) {
        // moments relative to the pivot atom (see kabsch_moments), reference rows as float4
        const float* p0 = xf + aoff[0];
        const float pvx = p0[0], pvy = p0[1], pvz = p0[2];
        float sx = 0.f, sy = 0.f, sz = 0.f;
        float h[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) h[q] = 0.f;
#pragma unroll 1
        for (int k = 0; k < p.n_align; ++k) {
          const float* pk = xf + aoff[k];
          const float4 yk = ref4[k];
          const float px = pk[0] - pvx, py = pk[1] - pvy, pz = pk[2] - pvz;
          sx += px; sy += py; sz += pz;
          h[0] = fmaf(px, yk.x, h[0]); h[1] = fmaf(px, yk.y, h[1]); h[2] = fmaf(px, yk.z, h[2]);
          h[3] = fmaf(py, yk.x, h[3]); h[4] = fmaf(py, yk.y, h[4]); h[5] = fmaf(py, yk.z, h[5]);
          h[6] = fmaf(pz, yk.x, h[6]); h[7] = fmaf(pz, yk.y, h[7]); h[8] = fmaf(pz, yk.z, h[8]);
        }
        const float inv_n = 1.0f / (float)p.n_align;
        rg.c[0] = fmaf(sx, inv_n, pvx);
        rg.c[1] = fmaf(sy, inv_n, pvy);
        rg.c[2] = fmaf(sz, inv_n, pvz);
#pragma unroll
        for (int q = 0; q < 9; ++q) rg.H[q] = h[q];
        kabsch_rotation(rg);
      } else {
#pragma unroll
        for (int q = 0; q < 9; ++q) rg.R[q] = (q == 0 || q == 4 || q == 8) ? 1.f : 0.f;
        rg.c[0] = rg.c[1] = rg.c[2] = 0.f;
      }
      const int ab = i & 1;
      unsigned char* a1buf = smem + wl.a1s_off[ab];
      SmemFeatOut out{a1buf + ft * 16, a1buf + kp0 * (WS_F * 4) + ft * 16};
      WS_EVT(g, i, 1);
      mbar_wait_hint(&a1_empty[ab], (uint32_t)(((i >> 1) & 1) ^ 1));   // the MMA that read this buffer is complete
      WS_EVT(g, i, 2);
      int cdone = 0;
      if (n_lead > 0) cdone = ws_position_features(xf, ent, n_lead, rg, out);
      {                                            // zero whole K-chunks up to kp0 (padding; the interpreter overwrites)
        const int cz = mixed ? ((3 * n_lead + 3) & ~3) : cdone;
        for (int c = cz; c < kp0; c += 4) {
          *reinterpret_cast<uint4*>(out.hi_row + (c >> 2) * (WS_F * 16)) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(out.lo_row + (c >> 2) * (WS_F * 16)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      if (mixed) {
        for (int e = n_lead; e < p.n_entries; ++e) {
          const Entry en = load_entry(ent + ENTRY_INTS * e);
          feature_forward(en, xf, aligned, rg, p.use_angle, out);
        }
      }
      fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core's async proxy
      mbar_arrive(&a1_full[ab]);
      mbar_arrive(&x_empty[b]);
      WS_EVT(g, i, 3);
    }
  } else if (warp < WS_W_E2) {
    // ================= E1 =================
    if (nh == 2) {
      const float* bias = reinterpret_cast<const float*>(smem + lay.bias_off[0]);
      const int w = (warp - WS_W_E1) >> 2;       // this warpgroup takes local tiles w, w + NE, ...
      int i = w;
      for (long long tile = first + (long long)w * stride; tile < ntiles; tile += WS_NE * stride, i += WS_NE) {
        const int ab = i % wl.n_a2buf;
        const uint32_t par_e = (uint32_t)(((i / wl.n_a2buf) & 1) ^ 1);
        const uint32_t lane_ahi = lane_base + (uint32_t)wl.col_a2[ab];
        const uint32_t lane_alo = lane_ahi + (uint32_t)lay.kp[1];
        const int db = i % ndb;
        ws_hidden_epilogue<ACT>(lane_base + wl.col_d1[db], lane_ahi, lane_alo, bias, lay.np[0], &d1_full[db],
                                &d1_free[db], (uint32_t)((i / ndb) & 1), &a2_empty[ab], par_e, &a2_full[ab], i);
      }
    }
  } else if (warp < WS_W_PROD) {
    // ================= E2 =================
    const float* bias = reinterpret_cast<const float*>(smem + lay.bias_off[nh - 1]);
    const float* wlast = reinterpret_cast<const float*>(smem + lay.wlast_off);
    const float* blast = reinterpret_cast<const float*>(smem + lay.blast_off);
    const int kout = p.dims[nl];
    const int np = lay.np[nh - 1];
    const int ft = tid & (WS_F - 1);            // frame within the tile
    const int w = (warp - WS_W_E2) >> 2;
    int i = w;
    for (long long tile = first + (long long)w * stride; tile < ntiles; tile += WS_NE * stride, i += WS_NE) {
      const long long f_base = tile * (long long)WS_F;
      const bool valid = f_base + ft < L;
      float* yrow = y + (f_base + ft) * kout;
      const int db = i % ndb;
      ws_final_epilogue<ACT>(lane_base + wl.col_d2[db], bias, np, &d2_full[db], &d2_free[db], (uint32_t)((i / ndb) & 1),
                             wlast, blast, kout, yrow, valid, i);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
#ifdef MOLANN_WS_TRACE
  if (blockIdx.x == 0 && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g_ws_trace[7 * 64 * 8 + 2] = clock64();
    g_ws_trace[7 * 64 * 8 + 3] = (long long)gt;
  }
  if (tid == 0 && blockIdx.x < 256) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g_ws_life[blockIdx.x * 2 + 1] = (long long)gt;
  }
#endif
  if (warp == 0) tmem_dealloc(tbase, 512u);
}

}  // namespace molann
