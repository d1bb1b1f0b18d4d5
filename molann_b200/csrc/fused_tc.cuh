// fused_tc.cuh -- fused small-system kernels with the MLP contraction on the 5th-gen tensor cores.
//
// Tile = 128 frames = 128 threads = 128 TMEM lanes: thread t owns frame t of the tile end to end.
//   1. TMA bulk copy stages the tile's coordinates (one contiguous byte range of x) into smem;
//   2. thread-per-frame Kabsch (Jacobi on Horn's 4x4) + feature program -> feature row;
//   3. every feature value is split into TF32 hi/lo and written straight to TMEM as the A operand
//      (tcgen05.st, warp-uniform column) -- features and activations never touch shared memory; weights
//      (hi/lo, chunk-major K-major) sit in smem for the whole kernel;
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
  int xs_off;                        // coordinate tile [F][3n] (forward kernel)
  int bhi_off[MOLANN_MAX_LAYERS];    // per MMA layer: weights hi / lo, chunk-major [Kp/4][Np][4]
  int blo_off[MOLANN_MAX_LAYERS];
  int thi_off[MOLANN_MAX_LAYERS];    // transposed weights (backward), chunk-major [Np/4][Kp][4]
  int tlo_off[MOLANN_MAX_LAYERS];
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

// activation value h and derivative d = act'(v) from ONE exponential.  The derivative is formed from
// (e, r) directly -- d = 4 e r^2 for tanh -- so saturated units keep full relative accuracy, which
// 1 - h*h (cancellation on the rounded h) cannot give.
template <int ACT>
__device__ __forceinline__ void act_value_and_grad_t(float v, float& h, float& d) {
  constexpr int act = ACT;
  if (act == ACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v * 2.8853900817779268f, 126.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    h = fmaf(-2.0f, r, 1.0f);
    d = 4.0f * (e * r) * r;
  } else if (act == ACT_SIGMOID) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v * -1.4426950408889634f, 126.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    h = r;
    d = (e * r) * r;
  } else if (act == ACT_RELU) {
    h = fmaxf(v, 0.f);
    d = v > 0.f ? 1.0f : 0.f;
  } else {
    h = v;
    d = 1.0f;
  }
}

// the same for two pre-activations (v0 + b0, v1 + b1) with the fp32 arithmetic as packed f32x2 operations
// (SASS FFMA2 / FADD2 / FMUL2: one issue slot for two results -- this kernel is issue and latency bound)
template <int ACT>
__device__ __forceinline__ void act_value_and_grad_x2(float v0, float v1, float b0, float b1, float& h0, float& h1,
                                                      float& d0, float& d1) {
  constexpr int act = ACT;
  if (act == ACT_TANH || act == ACT_SIGMOID) {
    const float sc = act == ACT_TANH ? 2.8853900817779268f : -1.4426950408889634f;
    float s0, s1;
    f2_unpack(f2_mul(f2_add(f2_pack(v0, v1), f2_pack(b0, b1)), f2_pack(sc, sc)), s0, s1);
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(s0, 126.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(s1, 126.0f)));
    const unsigned long long e = f2_pack(e0, e1);
    float p0, p1;
    f2_unpack(f2_add(e, f2_pack(1.0f, 1.0f)), p0, p1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(p0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(p1));
    const unsigned long long r = f2_pack(r0, r1);
    const unsigned long long err = f2_mul(f2_mul(e, r), r);                    // e r^2
    if (act == ACT_TANH) {
      f2_unpack(f2_fma(f2_pack(-2.0f, -2.0f), r, f2_pack(1.0f, 1.0f)), h0, h1);
      f2_unpack(f2_mul(err, f2_pack(4.0f, 4.0f)), d0, d1);
    } else {
      h0 = r0;
      h1 = r1;
      f2_unpack(err, d0, d1);
    }
  } else {
    act_value_and_grad_t<ACT>(v0 + b0, h0, d0);
    act_value_and_grad_t<ACT>(v1 + b1, h1, d1);
  }
}

// round-to-nearest TF32 split: |x - hi - lo| <= 2^-22 |x| once the MMA truncates lo to TF32
__device__ __forceinline__ void split_tf32_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// Feature column writer: value -> TF32 hi/lo -> this thread's TMEM lane (column = feature index).
struct TmemFeatOut {
  uint32_t hi_addr, lo_addr;
  __device__ __forceinline__ void operator()(int col, float v) {
    uint32_t hi, lo;
    split_tf32_rn(v, hi, lo);
    tmem_st1(hi_addr + col, hi);
    tmem_st1(lo_addr + col, lo);
  }
};
// Feature-cotangent reader straight from the accumulator columns of this thread's TMEM lane.
struct TmemGIn {
  uint32_t addr;
  __device__ __forceinline__ float operator()(int col) const {
    const float v = tmem_ld1_nowait(addr + col);
    tmem_wait_ld();
    return v;
  }
  __device__ __forceinline__ void load2(int col, float& a, float& b) const {
    a = tmem_ld1_nowait(addr + col);
    b = tmem_ld1_nowait(addr + col + 1);
    tmem_wait_ld();
  }
  __device__ __forceinline__ void load3(int col, float& a, float& b, float& c) const {
    a = tmem_ld1_nowait(addr + col);
    b = tmem_ld1_nowait(addr + col + 1);
    c = tmem_ld1_nowait(addr + col + 2);
    tmem_wait_ld();
  }
};

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&r)[12], int at) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[at + i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float (&r)[12], int at) {
  uint32_t u[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) r[at + i] = __uint_as_float(u[i]);
}
// N consecutive columns (N = 12, 6 or 3: the columns of 4, 2 or 1 position entries) as the widest stores that tile it
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[N]) {
  int c = 0;
  if (N - c >= 8) {
    tmem_st8(taddr + c, *reinterpret_cast<const uint32_t(*)[8]>(&v[c]));
    c += 8;
  }
  if (N - c >= 4) {
    tmem_st4(taddr + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
    c += 4;
  }
  if (N - c >= 2) {
    tmem_st2(taddr + c, v[c], v[c + 1]);
    c += 2;
  }
  if (N - c >= 1) tmem_st1(taddr + c, v[c]);
}

// Leading position entries, forward: z = (x - c) R = x R - t (t = c R) for NA entries from entry e0, split into
// TF32 hi/lo and stored as 3 NA columns of this thread's A-operand lane.
template <int NA>
__device__ __forceinline__ void tc_position_group(const float* __restrict__ xf, const int* __restrict__ ent, int e0,
                                                  const float (&R)[9], float t0, float t1, float t2,
                                                  uint32_t lane_ahi, uint32_t lane_alo) {
  uint32_t hi[3 * NA], lo[3 * NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const float* p = xf + 3 * ent[ENTRY_INTS * (e0 + i) + 1];
    const float px = p[0], py = p[1], pz = p[2];
    split_tf32_rn(fmaf(px, R[0], fmaf(py, R[3], fmaf(pz, R[6], -t0))), hi[3 * i], lo[3 * i]);
    split_tf32_rn(fmaf(px, R[1], fmaf(py, R[4], fmaf(pz, R[7], -t1))), hi[3 * i + 1], lo[3 * i + 1]);
    split_tf32_rn(fmaf(px, R[2], fmaf(py, R[5], fmaf(pz, R[8], -t2))), hi[3 * i + 2], lo[3 * i + 2]);
  }
  tmem_st_cols<3 * NA>(lane_ahi + 3 * e0, hi);
  tmem_st_cols<3 * NA>(lane_alo + 3 * e0, lo);
}
__device__ __forceinline__ void tc_position_features(const float* __restrict__ xf, const int* __restrict__ ent,
                                                     int n_lead, const Rigid& rg, uint32_t lane_ahi, uint32_t lane_alo) {
  const float t0 = fmaf(rg.c[0], rg.R[0], fmaf(rg.c[1], rg.R[3], rg.c[2] * rg.R[6]));
  const float t1 = fmaf(rg.c[0], rg.R[1], fmaf(rg.c[1], rg.R[4], rg.c[2] * rg.R[7]));
  const float t2 = fmaf(rg.c[0], rg.R[2], fmaf(rg.c[1], rg.R[5], rg.c[2] * rg.R[8]));
  int e0 = 0;
#pragma unroll 1
  for (; e0 + 4 <= n_lead; e0 += 4) tc_position_group<4>(xf, ent, e0, rg.R, t0, t1, t2, lane_ahi, lane_alo);
  if (e0 + 2 <= n_lead) {
    tc_position_group<2>(xf, ent, e0, rg.R, t0, t1, t2, lane_ahi, lane_alo);
    e0 += 2;
  }
  if (e0 < n_lead) tc_position_group<1>(xf, ent, e0, rg.R, t0, t1, t2, lane_ahi, lane_alo);
}

// Leading position entries, backward (aligned model): the cotangent columns of NA entries are read from this
// thread's accumulator lane in one go; per entry  M += (x - c)^T g,  t = g R^T,  sg += t,  gx[a0] += t.
template <int NA, class Acc>
__device__ __forceinline__ void tc_position_group_bwd(const float* __restrict__ xf, const int* __restrict__ ent, int e0,
                                                      const Rigid& rg, uint32_t lane_d, Acc& acc, float (&M)[9],
                                                      float (&sg)[3]) {
  float g[12];
  if (NA == 4) {
    tmem_ld8_nowait(lane_d + 3 * e0, g, 0);
    tmem_ld4_nowait(lane_d + 3 * e0 + 8, g, 8);
  } else {
#pragma unroll
    for (int c = 0; c < 3 * NA; ++c) g[c] = tmem_ld1_nowait(lane_d + 3 * e0 + c);
  }
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int a0 = ent[ENTRY_INTS * (e0 + i) + 1];
    const float* p = xf + 3 * a0;
    const float dx = p[0] - rg.c[0], dy = p[1] - rg.c[1], dz = p[2] - rg.c[2];
    const float g0 = g[3 * i], g1 = g[3 * i + 1], g2 = g[3 * i + 2];
    M[0] = fmaf(dx, g0, M[0]); M[1] = fmaf(dx, g1, M[1]); M[2] = fmaf(dx, g2, M[2]);
    M[3] = fmaf(dy, g0, M[3]); M[4] = fmaf(dy, g1, M[4]); M[5] = fmaf(dy, g2, M[5]);
    M[6] = fmaf(dz, g0, M[6]); M[7] = fmaf(dz, g1, M[7]); M[8] = fmaf(dz, g2, M[8]);
    V3 t;
    rot_transpose_apply(rg, g0, g1, g2, t.x, t.y, t.z);
    sg[0] += t.x; sg[1] += t.y; sg[2] += t.z;
    acc(a0, t);
  }
}

// zero the first `kp` hi/lo columns of the A operand (padding columns must not hold stale activations)
__device__ __forceinline__ void zero_a_operand(uint32_t lane_addr, uint32_t col_hi, uint32_t col_lo, int kp) {
  uint32_t z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) z[i] = 0u;
  for (int c = 0; c < kp; c += 16) {
    tmem_st16(lane_addr + col_hi + c, z);
    tmem_st16(lane_addr + col_lo + c, z);
  }
  tmem_wait_st();
}

// Split weights of MMA layer k into chunk-major hi/lo smem operands (zero padded to [np][kp]).
__device__ __forceinline__ void stage_tc_weights(const float* __restrict__ Wg, const float* __restrict__ bg, int K,
                                                 int N, int kp, int np, unsigned char* bhi, unsigned char* blo,
                                                 float* bias, int tid, int nthreads, unsigned char* thi = nullptr,
                                                 unsigned char* tlo = nullptr) {
  for (int idx = tid; idx < np * kp; idx += nthreads) {
    const int n = idx / kp, k = idx - n * kp;
    const float w = (n < N && k < K) ? Wg[(long long)n * K + k] : 0.f;
    uint32_t hi, lo;
    split_tf32_rn(w, hi, lo);
    lo = (lo + 0x1000u) & 0xffffe000u;
    const uint32_t off = chunk_major_offset(n, k, np);
    *reinterpret_cast<uint32_t*>(bhi + off) = hi;
    *reinterpret_cast<uint32_t*>(blo + off) = lo;
    if (thi != nullptr) {            // W^T as a K-major operand: rows = inputs (kp), contraction = outputs
      const uint32_t toff = chunk_major_offset(k, n, kp);
      *reinterpret_cast<uint32_t*>(thi + toff) = hi;
      *reinterpret_cast<uint32_t*>(tlo + toff) = lo;
    }
  }
  for (int n = tid; n < np; n += nthreads) bias[n] = (n < N) ? bg[n] : 0.f;
}

// one lane of a converged warp
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}

// D[128 x np] = A[128 x kp] * B^T with the 3xTF32 expansion, then commit to `mbar`.
// Called by ONE WHOLE (converged) warp; the elected lane issues.  The TMEM base is passed through a shuffle so
// that ptxas knows it is warp-uniform: the address arithmetic then stays in uniform registers and the
// tcgen05.mma instructions issue back to back.  (Issued from a divergent `if (tid == 0)` every MMA cost ~14
// instructions / ~85 cycles of ELECT / R2UR.BROADCAST glue: tests/cuda/ws_trace.cu.)
__device__ __forceinline__ void issue_layer_mma(uint32_t tbase_any, uint32_t colA_hi, uint32_t colA_lo, uint32_t colD,
                                                const unsigned char* bhi, const unsigned char* blo, int kp, int np,
                                                void* mbar) {
  const uint32_t tbase = __shfl_sync(0xffffffffu, tbase_any, 0);
  const uint32_t leader = elect_one();
  const uint32_t idesc = idesc_tf32(TC_F, np);
  const uint32_t bhi_a = smem_u32(bhi), blo_a = smem_u32(blo);
  const uint32_t step = 2u * (uint32_t)np * 16u;        // two 16-byte K-chunks per MMA (K = 8)
  const uint32_t lbo = (uint32_t)np * 16u;
  for (int j = 0; j < kp / 8; ++j) {
    const uint64_t bh = smem_desc_kmajor(bhi_a + j * step, lbo, 128);
    const uint64_t bl = smem_desc_kmajor(blo_a + j * step, lbo, 128);
    if (leader) {
      mma_tf32_ts(tbase + colD, tbase + colA_lo + 8 * j, bh, idesc, j > 0);      // small terms first
      mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bl, idesc, 1);
      mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bh, idesc, 1);
    }
  }
  if (leader) mma_commit(mbar);
  __syncwarp();
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
    // ---- geometry: thread t <-> frame t; features go straight to the A operand in TMEM ----
    {
      const int f = tid < nf ? tid : nf - 1;
      const float* xf = xs + f * n3;
      Rigid rg;
      const bool aligned = p.n_align > 0;
      if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      __syncwarp();
      zero_a_operand(lane_addr, COL_AHI, COL_ALO, lay.kp[0]);
      TmemFeatOut out{lane_addr + COL_AHI, lane_addr + COL_ALO};
      for (int e = 0; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    const long long next = tile + gridDim.x;
    if (next < ntiles && is_tma_tile(next)) issue_x(next);        // xs is free: overlap with the MLP
    if (warp == 0) {
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
        if (warp == 0) {
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

// =============================================================================================
// Value and gradient (forward recompute + d<gy,y>/dx) -- TILES frame tiles per CTA
// =============================================================================================
// Each warpgroup (4 warps = 128 threads) owns one 128-frame tile slot end to end and synchronises only
// with itself (named barrier + its own mbarriers), so the warpgroups of a CTA run phase-shifted and one
// group's tensor-core work overlaps the other's CUDA-core work while they share one smem copy of the
// weights (forward operand W and backward operand W^T, both K-major: an MN-major TF32 operand would need
// the 32B-base 128B swizzle, which no K-major layout of the same bytes matches) and ONE gradient staging
// tile, handed over with a lock (a warpgroup holds it only from zero-fill to the end of its bulk store).
struct TcVgLayout {
  TcLayout base;                       // shared: weights, biases, plan constants
  int xs_off[4];                       // per tile slot: coordinate tile
  int gxs_off;                         // shared gradient tile
  int lock_off;
  int mbar_off, tptr_off;
  int total_bytes;
};

__device__ __forceinline__ void wg_sync(int wg) {
  asm volatile("bar.sync %0, %1;" ::"r"(wg + 1), "r"(128) : "memory");
}

#ifdef MOLANN_WS_TRACE
__device__ long long g_vg_trace[64 * 16];
__device__ long long g_vg_life[256 * 4];       // per CTA: start, end of warpgroup 0, end of warpgroup 1 (globaltimer ns)
__device__ __forceinline__ long long vg_gtime() {
  unsigned long long gt;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
  return (long long)gt;
}
#define VG_EVT(it, ev)                                                                       \
  do {                                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (it) < 64) g_vg_trace[(it) * 16 + (ev)] = clock64(); \
  } while (0)
#else
#define VG_EVT(it, ev) \
  do {                 \
  } while (0)
#endif

// ACT is a template parameter: with the activation chosen at run time every element carried its own branch, the
// 16 elements of a chunk could not overlap, and the two epilogues took 8k + 11k of a tile's 42k cycles
// (tests/cuda/ws_trace.cu).
// JAC: Jacobian mode (value_and_jacobian): no cotangent; the forward runs once per tile, act'(z) of
// BOTH hidden layers stays parked in TMEM (64 more columns: needs TILES == 1), and the backward chain -- two small
// contractions, feature + alignment backward, tile store -- runs once per output with gz = W_last[o, :] * act'(z),
// writing plane o of gx[kout][L][n_inp][3] (plane-major, so a tile of a plane is one contiguous bulk store).
template <int TILES, int ACT, bool JAC = false>
__global__ void __launch_bounds__(TILES * TC_F, 1)
fused_tc_value_grad_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ TcVgLayout vl,
                           const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ y,
                           float* __restrict__ gx, long long L, int use_tma) {
  static_assert(!JAC || TILES == 1, "Jacobian mode keeps five 64-column blocks per tile in TMEM");
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int NT = TILES * TC_F;
  const TcLayout& lay = vl.base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int wg = tid >> 7, wt = tid & 127;
  const int n3 = 3 * p.n_inp;
  const int nl = p.n_layers;
  const int nh = nl - 1;               // hidden (tensor-core) layers: 1 or 2
  float* xs = reinterpret_cast<float*>(smem + vl.xs_off[wg]);
  float* gxs = reinterpret_cast<float*>(smem + vl.gxs_off);
  int* gx_lock = reinterpret_cast<int*>(smem + vl.lock_off);
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  unsigned long long* mbar_x = reinterpret_cast<unsigned long long*>(smem + vl.mbar_off) + 2 * wg;
  unsigned long long* mbar_mma = mbar_x + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + vl.tptr_off);

  stage_tc_consts<NT>(p, lay, smem, tid);
  for (int k = 0; k < nh; ++k)
    stage_tc_weights(p.W[k], p.b[k], p.dims[k], p.dims[k + 1], lay.kp[k], lay.np[k], smem + lay.bhi_off[k],
                     smem + lay.blo_off[k], reinterpret_cast<float*>(smem + lay.bias_off[k]), tid, NT,
                     smem + lay.thi_off[k], smem + lay.tlo_off[k]);
  {
    const int K = p.dims[nl - 1], N = p.dims[nl];
    float* wl = reinterpret_cast<float*>(smem + lay.wlast_off);
    float* bl = reinterpret_cast<float*>(smem + lay.blast_off);
    for (int i = tid; i < N * TC_MAXW; i += NT) {
      const int o = i / TC_MAXW, j = i - o * TC_MAXW;
      wl[i] = (j < K) ? p.W[nl - 1][(long long)o * K + j] : 0.f;
    }
    for (int o = tid; o < N; o += NT) bl[o] = p.b[nl - 1][o];
  }
  if (tid == 0) *gx_lock = 0;
  if (wt == 0) {
    mbar_init(mbar_x, 1);
    mbar_init(mbar_mma, 1);
    fence_mbar_init();
  }
  constexpr uint32_t TMEM_COLS = JAC ? 512u : (uint32_t)TILES * 256u;
  if (warp == 0) tmem_alloc(tptr, TMEM_COLS);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tptr + (uint32_t)wg * 256u;
  const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr uint32_t COL_AHI = 0, COL_ALO = 64, COL_D = 128, COL_H1 = 192, COL_H2 = 256;

  const long long ntiles = (L + TC_F - 1) / TC_F;
  const uint32_t tile_bytes = (uint32_t)TC_F * (uint32_t)n3 * 4u;
  const int kout = p.dims[nl];
  uint32_t phase_x = 0, phase_m = 0;
  auto is_tma_tile = [&](long long t) { return use_tma && (t + 1) * (long long)TC_F <= L; };
  auto issue_x = [&](long long t) {
    if (wt == 0) {
      mbar_expect_tx(mbar_x, tile_bytes);
      bulk_g2s(xs, x + t * (long long)TC_F * n3, tile_bytes, mbar_x);
    }
  };
  auto wait_mma = [&]() {
    mbar_wait(mbar_mma, phase_m);
    phase_m ^= 1u;
    tc_fence_after_sync();
    __syncwarp();
  };
  auto publish_a_and_issue = [&](const unsigned char* bh, const unsigned char* bl, int kp_, int np_) {
    tmem_wait_st();
    tc_fence_before_sync();
    wg_sync(wg);
    if ((wt >> 5) == 0) {                 // first warp of the warpgroup, converged
      tc_fence_after_sync();
      issue_layer_mma(tbase, COL_AHI, COL_ALO, COL_D, bh, bl, kp_, np_, mbar_mma);
    }
  };
  const long long tstride = (long long)gridDim.x * TILES;
  long long tile = (long long)blockIdx.x * TILES + wg;
  if (tile < ntiles && is_tma_tile(tile)) issue_x(tile);

  int n_lead = 0;                              // leading position entries take the unrolled paths
  while (n_lead < p.n_entries && ent[ENTRY_INTS * n_lead] == FEAT_POSITION) ++n_lead;
  const bool mixed = n_lead < p.n_entries;
#ifdef MOLANN_WS_TRACE
  if (tid == 0 && blockIdx.x < 256) g_vg_life[blockIdx.x * 4 + 0] = vg_gtime();
#endif
  int vg_it = -1;
  for (; tile < ntiles; tile += tstride) {
    ++vg_it;
    VG_EVT(vg_it, 0);
    const long long f_base = tile * (long long)TC_F;
    const int nf = (int)((L - f_base) < (long long)TC_F ? (L - f_base) : (long long)TC_F);
    if (is_tma_tile(tile)) {
      mbar_wait(mbar_x, phase_x);
      phase_x ^= 1u;
    } else {
      const float* src = x + f_base * n3;
      for (int i = wt; i < nf * n3; i += TC_F) xs[i] = src[i];
      wg_sync(wg);
    }
    VG_EVT(vg_it, 1);
    float go[8];                               // this frame's output cotangent: issued now, consumed after two MMAs
#pragma unroll
    for (int o = 0; o < 8; ++o) go[o] = (!JAC && o < kout && wt < nf) ? __ldg(gy + (f_base + wt) * kout + o) : 0.f;
    // ---- geometry ----
    const int f = wt < nf ? wt : nf - 1;
    const float* xf = xs + f * n3;
    Rigid rg;
    const bool aligned = p.n_align > 0;
    if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
    __syncwarp();
    VG_EVT(vg_it, 2);
    if (!aligned) {
#pragma unroll
      for (int q = 0; q < 9; ++q) rg.R[q] = (q == 0 || q == 4 || q == 8) ? 1.f : 0.f;
      rg.c[0] = rg.c[1] = rg.c[2] = 0.f;
    }
    if (mixed) {
      zero_a_operand(lane_addr, COL_AHI, COL_ALO, lay.kp[0]);
    } else {                                   // only the padding columns [d_feat, kp) need zeros
      for (int c = p.d_feat; c < lay.kp[0]; ++c) {
        tmem_st1(lane_addr + COL_AHI + c, 0u);
        tmem_st1(lane_addr + COL_ALO + c, 0u);
      }
    }
    if (n_lead > 0) tc_position_features(xf, ent, n_lead, rg, lane_addr + COL_AHI, lane_addr + COL_ALO);
    if (mixed) {
      TmemFeatOut out{lane_addr + COL_AHI, lane_addr + COL_ALO};
      for (int e = n_lead; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    VG_EVT(vg_it, 3);
    publish_a_and_issue(smem + lay.bhi_off[0], smem + lay.blo_off[0], lay.kp[0], lay.np[0]);
    // The epilogues are ROLLED 16-column loops: this kernel's executed path was ~70 KB of straight-line code per
    // tile against a 32 KB L1.5 instruction cache ("no instruction" = 20 % of its stalls, profiles/r1_g).
    // ---- first hidden layer (only when there are two): h_1 -> A operand, act'(z_1) parked in TMEM ----
    if (nh == 2) {
      wait_mma();
      VG_EVT(vg_it, 4);
      const float* bias = reinterpret_cast<const float*>(smem + lay.bias_off[0]);
      const int np = lay.np[0];
#pragma unroll 1
      for (int c = 0; c < np; c += 16) {
        float z[16];
        tmem_ld16(lane_addr + COL_D + c, z);
        tmem_wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(bias + c);
        uint32_t dv[16], hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b = b4[q];
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            float hh0, hh1, dd0, dd1;
            act_value_and_grad_x2<ACT>(z[4 * q + i], z[4 * q + i + 1], bb[i], bb[i + 1], hh0, hh1, dd0, dd1);
            dv[4 * q + i] = __float_as_uint(dd0);
            dv[4 * q + i + 1] = __float_as_uint(dd1);
            split_tf32_rn_x2(hh0, hh1, hi[4 * q + i], hi[4 * q + i + 1], lo[4 * q + i], lo[4 * q + i + 1]);
          }
        }
        tmem_st16(lane_addr + COL_H1 + c, dv);
        tmem_st16(lane_addr + COL_AHI + c, hi);
        tmem_st16(lane_addr + COL_ALO + c, lo);
      }
      VG_EVT(vg_it, 5);
      publish_a_and_issue(smem + lay.bhi_off[1], smem + lay.blo_off[1], lay.kp[1], lay.np[1]);
    }
    // ---- last hidden layer, 16 columns at a time: y += h W_last^T, gz = (gy W_last) * act'(z) ----
    {
      const float* wl = reinterpret_cast<const float*>(smem + lay.wlast_off);
      const float* bl = reinterpret_cast<const float*>(smem + lay.blast_off);
      const float* bias = reinterpret_cast<const float*>(smem + lay.bias_off[nh - 1]);
      const int np = lay.np[nh - 1];
      float yacc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) yacc[o] = (o < kout) ? bl[o] : 0.f;
      VG_EVT(vg_it, 6);
      wait_mma();
      VG_EVT(vg_it, 7);
#pragma unroll 1
      for (int c = 0; c < np; c += 16) {
        float z[16];
        tmem_ld16(lane_addr + COL_D + c, z);
        float gh[16];                              // (gy W_last)[c .. c+15], formed while the load is in flight
#pragma unroll
        for (int i = 0; i < 16; ++i) gh[i] = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (o < kout) {
            const float4* w4 = reinterpret_cast<const float4*>(wl + o * TC_MAXW + c);
            const float g_o = go[o];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 w = w4[q];
              gh[4 * q] = fmaf(g_o, w.x, gh[4 * q]);
              gh[4 * q + 1] = fmaf(g_o, w.y, gh[4 * q + 1]);
              gh[4 * q + 2] = fmaf(g_o, w.z, gh[4 * q + 2]);
              gh[4 * q + 3] = fmaf(g_o, w.w, gh[4 * q + 3]);
            }
          }
        }
        tmem_wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(bias + c);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b = b4[q];
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            const int c0i = 4 * q + i;
            float hh0, hh1, dd0, dd1;
            act_value_and_grad_x2<ACT>(z[c0i], z[c0i + 1], bb[i], bb[i + 1], hh0, hh1, dd0, dd1);
            z[c0i] = hh0;
            z[c0i + 1] = hh1;
            if (JAC) {                                                            // parked: every plane needs it
              hi[c0i] = __float_as_uint(dd0);
              hi[c0i + 1] = __float_as_uint(dd1);
            } else {                                                              // gz of the last hidden layer
              float g0, g1;
              f2_unpack(f2_mul(f2_pack(gh[c0i], gh[c0i + 1]), f2_pack(dd0, dd1)), g0, g1);
              split_tf32_rn_x2(g0, g1, hi[c0i], hi[c0i + 1], lo[c0i], lo[c0i + 1]);
            }
          }
        }
        if (JAC) {
          tmem_st16(lane_addr + COL_H2 + c, hi);
        } else {
          tmem_st16(lane_addr + COL_AHI + c, hi);
          tmem_st16(lane_addr + COL_ALO + c, lo);
        }
        if (y != nullptr) {
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            if (o < kout) {
              const float4* w4 = reinterpret_cast<const float4*>(wl + o * TC_MAXW + c);
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w = w4[q];
                a = fmaf(z[4 * q], w.x, a);
                b = fmaf(z[4 * q + 1], w.y, b);
                a = fmaf(z[4 * q + 2], w.z, a);
                b = fmaf(z[4 * q + 3], w.w, b);
              }
              yacc[o] += a + b;
            }
          }
        }
      }
      if (y != nullptr && wt < nf) {
#pragma unroll
        for (int o = 0; o < 8; ++o)
          if (o < kout) y[(f_base + wt) * kout + o] = yacc[o];
      }
    }
    VG_EVT(vg_it, 8);
    const int nplanes = JAC ? kout : 1;
#pragma unroll 1
    for (int plane = 0; plane < nplanes; ++plane) {
    if (JAC) {                   // gz of the last hidden layer for output `plane`: W_last[plane, :] * act'(z)
      const float* wl = reinterpret_cast<const float*>(smem + lay.wlast_off) + plane * TC_MAXW;
      const int np = lay.np[nh - 1];
#pragma unroll 1
      for (int c = 0; c < np; c += 16) {
        float dk[16];
        tmem_ld16(lane_addr + COL_H2 + c, dk);
        tmem_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_tf32_rn(wl[c + i] * dk[i], hi[i], lo[i]);
        tmem_st16(lane_addr + COL_AHI + c, hi);
        tmem_st16(lane_addr + COL_ALO + c, lo);
      }
    }
    // ---- backward through the tensor-core layers: gh_k = gz_{k+1} W_k  (operand W^T, K-major) ----
    for (int k = nh - 1; k >= 0; --k) {
      const int kb = lay.np[k];            // contraction width (outputs of forward layer k)
      const int nb = lay.kp[k];            // result width (inputs of forward layer k)
      publish_a_and_issue(smem + lay.thi_off[k], smem + lay.tlo_off[k], kb, nb);
      wait_mma();
      if (k > 0) {               // gz_k = gh_k * act'(z_k) (parked in TMEM) -> A operand of the next contraction
#pragma unroll 1
        for (int c = 0; c < nb; c += 16) {
          float gh[16], dk[16];
          tmem_ld16(lane_addr + COL_D + c, gh);
          tmem_ld16(lane_addr + COL_H1 + c, dk);
          tmem_wait_ld();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float g0, g1;
            f2_unpack(f2_mul(f2_pack(gh[i], gh[i + 1]), f2_pack(dk[i], dk[i + 1])), g0, g1);
            split_tf32_rn_x2(g0, g1, hi[i], hi[i + 1], lo[i], lo[i + 1]);
          }
          tmem_st16(lane_addr + COL_AHI + c, hi);
          tmem_st16(lane_addr + COL_ALO + c, lo);
        }
      }
    }
    VG_EVT(vg_it, 9);
    // the feature cotangent now sits in accumulator columns [COL_D, COL_D + d_feat) of this lane
    if (TILES > 1) {             // take the shared gradient tile
      if (wt == 0)
        while (atomicCAS(gx_lock, 0, 1) != 0) __nanosleep(32);
      wg_sync(wg);
    }
    VG_EVT(vg_it, 10);
    {
      float4* g4 = reinterpret_cast<float4*>(gxs);       // tile bytes are a multiple of 16
      for (int i = wt; i < (TC_F * n3) / 4; i += TC_F) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    wg_sync(wg);
    VG_EVT(vg_it, 11);
    {
      TmemGIn gin{lane_addr + COL_D};
      RowAcc acc{gxs + wt * n3};
      float M[9], sg[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) M[i] = 0.f;
      sg[0] = sg[1] = sg[2] = 0.f;
      int e_first = 0;
      if (aligned && n_lead > 0) {             // unrolled position entries
        const uint32_t lane_d = lane_addr + COL_D;
        int e0 = 0;
#pragma unroll 1
        for (; e0 + 4 <= n_lead; e0 += 4) tc_position_group_bwd<4>(xf, ent, e0, rg, lane_d, acc, M, sg);
        if (e0 + 2 <= n_lead) {
          tc_position_group_bwd<2>(xf, ent, e0, rg, lane_d, acc, M, sg);
          e0 += 2;
        }
        if (e0 < n_lead) tc_position_group_bwd<1>(xf, ent, e0, rg, lane_d, acc, M, sg);
        e_first = n_lead;
      }
      for (int e = e_first; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_backward(en, xf, aligned, rg, p.use_angle, gin, acc, M, sg);
      }
      if (aligned) {
        float dH[9];
        align_backward_dH(rg, M, dH);
        const float inv_na = 1.0f / (float)p.n_align;
        for (int k = 0; k < p.n_align; ++k)
          acc(aidx[k], align_atom_grad(dH, sg, inv_na, ref[3 * k], ref[3 * k + 1], ref[3 * k + 2]));
      }
    }
    VG_EVT(vg_it, 12);
    fence_proxy_async_smem();
    wg_sync(wg);
    const long long next = tile + tstride;
    if (plane == nplanes - 1 && next < ntiles && is_tma_tile(next)) issue_x(next);       // xs is free
    float* dst = gx + ((long long)plane * L + f_base) * n3;
    if (is_tma_tile(tile)) {
      if (wt == 0) {
        bulk_s2g(dst, gxs, tile_bytes);
        bulk_commit();
        bulk_wait_read0();                                       // smem source fully read
        if (TILES > 1) {
          __threadfence_block();
          atomicExch(gx_lock, 0);
        }
      }
      if (TILES == 1) wg_sync(wg);                               // gxs reusable by this group's next tile
    } else {
      for (int i = wt; i < nf * n3; i += TC_F) dst[i] = gxs[i];
      wg_sync(wg);
      if (TILES > 1 && wt == 0) {
        __threadfence_block();
        atomicExch(gx_lock, 0);
      }
    }
    }                            // planes
  }
  if (wt == 0) bulk_wait0();
#ifdef MOLANN_WS_TRACE
  if (wt == 0 && blockIdx.x < 256) g_vg_life[blockIdx.x * 4 + 1 + wg] = vg_gtime();
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tptr, TMEM_COLS);
}

}  // namespace molann
