// gemm_tc.cuh -- wide dense layers of the general path on the tensor cores:
//     C[M x N] = epilogue( A[M x K] * B[N x K]^T )         fp32 in, fp32 out, 3xTF32 inside (fp32-accurate)
// Used for the MLP layers that are too wide for the fused small-system kernels (C3: 800 -> 256 -> 128, C5:
// 2000 -> 256 -> 128) and for the matching backward contractions gz * W.  Replaces the FFMA tile kernel
// (general.cuh gemm_kernel) there; that kernel stays for parameter gradients and narrow layers.
//
// Structure (persistent CTA per SM, 20 warps, roles over mbarriers -- the same machinery as fused_ws.cuh):
//   producer  (1 warp)   cp.async.bulk of the pre-packed weight block of one 32-wide K-chunk (hi + lo contiguous)
//   converter (4 warps)  thread = row of the 128-row tile: 32 floats of its row -> TF32 hi / lo -> canonical
//                        K-major smem tiles (16-byte stores, conflict free); the next TWO chunks of the row are
//                        always in flight in registers (the A stream comes from HBM) and the lines after that are
//                        prefetched into L2
//   MMA       (1 warp)   elected lane: 4 K-steps x 3 tcgen05.mma (SS form) per chunk, accumulating in TMEM
//                        (N up to 256 -> 128 cycles per MMA: the tensor pipe runs at its full TF32 rate)
//   epilogue  (4 x 4)    thread = row x 64 columns.  The tensor core's fp32 accumulator rounds toward zero on every
//                        MMA, a bias that grows with the number of accumulation steps (K = 800 -> 300 steps -> 1e-5
//                        relative, measured); so K is cut into segments of 128 that start from a zeroed TMEM
//                        accumulator (double-buffered, 2 x 256 columns) and are summed in fp32 registers with
//                        round-to-nearest by these warps, which then apply bias / activation / activation
//                        gradient and write the row.
// Weights are packed once per call by gemm_tc_pack_kernel into per-(N-tile, K-chunk) blocks in the layout the MMA
// wants, so a block is one contiguous bulk copy.
#pragma once
#include "common.cuh"
#include "fused_tc.cuh"
#include "fused_ws.cuh"
#include "geometry.cuh"
#include "tc.cuh"

namespace molann {

constexpr int GT_M = 128;                      // rows per tile
constexpr int GT_KC = 32;                      // K per chunk (8 K-major 16-byte chunks)
constexpr int GT_NMAX = 256;                   // columns per tile
constexpr int GT_SEG = 1;                      // default K-chunks per TMEM accumulation segment (32 K = 12 MMAs)
constexpr int GT_THREADS = 28 * 32;            // 4 converter + 16 epilogue + (producer, MMA, 2 idle) + 4 idle warps
// setmaxnreg budgets.  The pool is what the CTA was LAUNCHED with (896 threads x 72 registers = 64512), not the SM's
// register file: increases beyond it wait forever.  4*96 + 16*88 + 4*32 + 4*24 warps x 32 = 64512.  The idle
// warpgroup exists only to donate its registers.
constexpr int GT_REGS_CONV = 96, GT_REGS_EPI = 88, GT_REGS_CTRL = 32, GT_REGS_IDLE = 24;
constexpr int GT_A_BYTES = GT_M * GT_KC * 4;   // one of hi / lo
constexpr int GT_STAGE_BYTES = 2 * GT_A_BYTES + 2 * GT_NMAX * GT_KC * 4;   // 96 KB
constexpr int GT_SMEM_BYTES = 2 * GT_STAGE_BYTES + 256;

enum { GT_EPI_BIAS_ACT = 0, GT_EPI_DACT = 1 };

__host__ __device__ inline int gt_np(int N, int n_tile) {      // padded column count of N-tile `n_tile`
  const int rest = N - n_tile * GT_NMAX;
  return round_up(rest < GT_NMAX ? rest : GT_NMAX, 16);
}
__host__ __device__ inline long long gt_pack_floats(int N, int K) {
  const int nkc = (K + GT_KC - 1) / GT_KC;
  const int nt = (N + GT_NMAX - 1) / GT_NMAX;
  return (long long)nt * nkc * 2 * GT_NMAX * GT_KC;             // every tile slot sized for a full tile
}

// B[n][k] = W[n * rs + k * cs] -> blocks (n_tile, k_chunk): hi [(k/4)][np][4] then lo, zero padded
__global__ void gemm_tc_pack_kernel(const float* __restrict__ W, long long rs, long long cs, int N, int K,
                                    float* __restrict__ Bp) {
  const int nkc = (K + GT_KC - 1) / GT_KC;
  const int Kp = nkc * GT_KC;
  const int nt = (N + GT_NMAX - 1) / GT_NMAX;
  const long long total = (long long)nt * GT_NMAX * Kp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e / ((long long)GT_NMAX * Kp));
    const int r = (int)(e - (long long)t * GT_NMAX * Kp);
    const int n = r / Kp, k = r - n * Kp;
    const int np = gt_np(N, t);
    if (n >= np) continue;
    const int ng = t * GT_NMAX + n;
    const float w = (ng < N && k < K) ? __ldg(W + (long long)ng * rs + (long long)k * cs) : 0.f;
    uint32_t hi, lo;
    split_tf32_rn(w, hi, lo);
    lo = (lo + 0x1000u) & 0xffffe000u;
    const int kc = k / GT_KC, kk = k - kc * GT_KC;
    float* blk = Bp + ((long long)t * nkc + kc) * (2 * GT_NMAX * GT_KC);
    const int off = ((kk >> 2) * np + n) * 4 + (kk & 3);
    blk[off] = __uint_as_float(hi);
    blk[np * GT_KC + off] = __uint_as_float(lo);
  }
}

// cursor over the (work item, K-chunk) pairs of this CTA, in the order every role visits them
struct GtCursor {
  long long item;
  int kc;
  __device__ __forceinline__ void step(int by, int nkc, long long stride) {
    kc += by;
    while (kc >= nkc) {
      kc -= nkc;
      item += stride;
    }
  }
};

// 32 consecutive floats of one A row (zeros outside the matrix)
__device__ __forceinline__ void gt_load_row_chunk(const float* __restrict__ A, long long lda, long long M, int K,
                                                  long long row, int k0, bool vec, float4 (&v)[8]) {
  if (row < M && vec && k0 + GT_KC <= K) {
    const float4* src = reinterpret_cast<const float4*>(A + row * lda + k0);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(src + q);
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + 4 * q + i;
        t[i] = (row < M && k < K) ? __ldg(A + row * lda + k) : 0.f;
      }
      v[q] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
}
// ... split (round to nearest: a truncating split leaves a remainder with the sign of x that the tensor core
// truncates again, a bias that does not average out over a long K) and stored as this row's 16-byte K-chunks
__device__ __forceinline__ void gt_store_row_chunk(unsigned char* a_hi, const float4 (&v)[8]) {
  unsigned char* a_lo = a_hi + GT_A_BYTES;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_tf32_rn(v[q].x, h0, l0);
    split_tf32_rn(v[q].y, h1, l1);
    split_tf32_rn(v[q].z, h2, l2);
    split_tf32_rn(v[q].w, h3, l3);
    *reinterpret_cast<uint4*>(a_hi + q * (GT_M * 16)) = make_uint4(h0, h1, h2, h3);
    *reinterpret_cast<uint4*>(a_lo + q * (GT_M * 16)) = make_uint4(l0, l1, l2, l3);
  }
}

template <int EPI>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const float* __restrict__ A, long long lda, long long M, int K, const float* __restrict__ Bp, int N,
               float* __restrict__ C, long long ldc, const float* __restrict__ bias, const float* __restrict__ H,
               int act, int apply_act, int seg) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 2 * GT_STAGE_BYTES);
  unsigned long long* empty = bars;            // [2] stage free (MMAs that read it are complete)
  unsigned long long* a_full = bars + 2;       // [2]
  unsigned long long* b_full = bars + 4;       // [2]
  unsigned long long* d_full = bars + 6;       // [2]
  unsigned long long* d_free = bars + 8;       // [2]
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 10);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&empty[s], 1);
      mbar_init(&a_full[s], 128);
      mbar_init(&b_full[s], 1);
      mbar_init(&d_full[s], 1);
      mbar_init(&d_free[s], 512);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tptr, 512u);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tptr != 0u) __trap();                   // full allocation: base 0 (keeps tcgen05 addresses warp-uniform)

  const int nkc = (K + GT_KC - 1) / GT_KC;
  const int nt = (N + GT_NMAX - 1) / GT_NMAX;
  const long long mt = (M + GT_M - 1) / GT_M;
  const long long nitems = mt * nt;
  const long long first = blockIdx.x, stride = gridDim.x;
  // every CTA walks the K-chunks from a different starting point (order is irrelevant to the sum): all 148 CTAs
  // asking the L2 for the same 64 KB weight block at the same moment was the bottleneck of the first version
  const int krot = (int)((blockIdx.x * 7u) % (unsigned)nkc);

  if (warp >= 24) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GT_REGS_IDLE));
  else if (warp >= 20) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GT_REGS_CTRL));

  if (warp == 20) {
    // ================= producer: packed weight blocks =================
    if (lane == 0) {
      int g = 0;
      for (long long item = first; item < nitems; item += stride) {
        const int n_tile = (int)(item % nt);
        const int np = gt_np(N, n_tile);
        const uint32_t bytes = 2u * (uint32_t)np * GT_KC * 4u;
        const float* src = Bp + (long long)n_tile * nkc * (2 * GT_NMAX * GT_KC);
        for (int kc = 0; kc < nkc; ++kc, ++g) {
          const int s = g & 1;
          mbar_wait_hint(&empty[s], (uint32_t)(((g >> 1) & 1) ^ 1));
          mbar_expect_tx(&b_full[s], bytes);
          bulk_g2s(smem + s * GT_STAGE_BYTES + 2 * GT_A_BYTES, src + (long long)((kc + krot) % nkc) * (2 * GT_NMAX * GT_KC), bytes,
                   &b_full[s]);
        }
      }
    }
  } else if (warp == 21) {
    // ================= MMA issuer =================
    const uint32_t leader = elect_one();
    int g = 0, sg = 0;                          // chunk and segment counters of this CTA
    for (long long item = first; item < nitems; item += stride) {
      const int n_tile = (int)(item % nt);
      const int np = gt_np(N, n_tile);
      const uint32_t idesc = idesc_tf32(GT_M, np);
      const uint32_t lbo_b = (uint32_t)np * 16u;
      for (int kc = 0; kc < nkc; ++kc, ++g) {
        const int db = sg & 1;
        const uint32_t d = (uint32_t)db * GT_NMAX;
        const bool seg_first = (kc % seg) == 0;
        const bool seg_last = (kc % seg) == seg - 1 || kc == nkc - 1;
        if (seg_first) {
          mbar_wait_hint(&d_free[db], (uint32_t)(((sg >> 1) & 1) ^ 1));
          tc_fence_after_sync();
        }
        const int s = g & 1;
        const uint32_t par = (uint32_t)((g >> 1) & 1);
        mbar_wait_hint(&a_full[s], par);
        mbar_wait_hint(&b_full[s], par);
        tc_fence_after_sync();
        const uint32_t a_hi = smem_u32(smem + s * GT_STAGE_BYTES), a_lo = a_hi + GT_A_BYTES;
        const uint32_t b_hi = a_hi + 2 * GT_A_BYTES, b_lo = b_hi + (uint32_t)np * GT_KC * 4u;
        // the chunk's eight cross-term MMAs first, its four leading-term MMAs last: every accumulation step
        // truncates at the magnitude the accumulator has reached, and in a fresh accumulator (seg = 1: every
        // chunk) the cross terms are 2^-11 of the result, so only the four leading steps round at full scale
#pragma unroll 1
        for (int j = 0; j < GT_KC / 8; ++j) {
          const uint64_t ah = smem_desc_kmajor(a_hi + j * (2u * GT_M * 16u), GT_M * 16u, 128);
          const uint64_t al = smem_desc_kmajor(a_lo + j * (2u * GT_M * 16u), GT_M * 16u, 128);
          const uint64_t bh = smem_desc_kmajor(b_hi + j * (2u * lbo_b), lbo_b, 128);
          const uint64_t bl = smem_desc_kmajor(b_lo + j * (2u * lbo_b), lbo_b, 128);
          if (leader) {
            mma_tf32_ss(d, al, bh, idesc, (!seg_first || j > 0) ? 1u : 0u);
            mma_tf32_ss(d, ah, bl, idesc, 1);
          }
        }
#pragma unroll 1
        for (int j = 0; j < GT_KC / 8; ++j) {
          const uint64_t ah = smem_desc_kmajor(a_hi + j * (2u * GT_M * 16u), GT_M * 16u, 128);
          const uint64_t bh = smem_desc_kmajor(b_hi + j * (2u * lbo_b), lbo_b, 128);
          if (leader) mma_tf32_ss(d, ah, bh, idesc, 1);
        }
        if (leader) mma_commit(&empty[s]);
        if (seg_last) {
          if (leader) mma_commit(&d_full[db]);
          ++sg;
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ================= converter: A rows -> TF32 hi / lo operand tiles =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GT_REGS_CONV));
    const int r = tid & 127;
    const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
    GtCursor cur{first, 0};                    // chunk being written
    auto load_chunk = [&](const GtCursor& c, float4 (&v)[8]) {
      if (c.item < nitems)
        gt_load_row_chunk(A, lda, M, K, (c.item / nt) * GT_M + r, ((c.kc + krot) % nkc) * GT_KC, vec, v);
    };
    auto prefetch_l2 = [&](GtCursor c, int ahead) {
      c.step(ahead, nkc, stride);
      if (c.item >= nitems) return;
      const long long row = (c.item / nt) * GT_M + r;
      const int k0 = ((c.kc + krot) % nkc) * GT_KC;
      if (row < M && k0 < K) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + row * lda + k0));
    };
    auto store_chunk = [&](const float4 (&v)[8], int s) { gt_store_row_chunk(smem + s * GT_STAGE_BYTES + r * 16, v); };
    float4 v0[8], v1[8];
    GtCursor c1 = cur;
    c1.step(1, nkc, stride);
    load_chunk(cur, v0);
    load_chunk(c1, v1);
    for (int a = 2; a < 6; ++a) prefetch_l2(cur, a);
    int g = 0;
    while (cur.item < nitems) {
      // even chunk (registers v0, stage 0), then odd chunk (v1, stage 1); each is refilled two chunks ahead
      mbar_wait_hint(&empty[0], (uint32_t)(((g >> 1) & 1) ^ 1));
      store_chunk(v0, 0);
      fence_proxy_async_smem();
      mbar_arrive(&a_full[0]);
      {
        GtCursor c2 = cur;
        c2.step(2, nkc, stride);
        load_chunk(c2, v0);
        prefetch_l2(cur, 6);
      }
      cur.step(1, nkc, stride);
      ++g;
      if (cur.item >= nitems) break;
      mbar_wait_hint(&empty[1], (uint32_t)(((g >> 1) & 1) ^ 1));
      store_chunk(v1, 1);
      fence_proxy_async_smem();
      mbar_arrive(&a_full[1]);
      {
        GtCursor c2 = cur;
        c2.step(2, nkc, stride);
        load_chunk(c2, v1);
        prefetch_l2(cur, 6);
      }
      cur.step(1, nkc, stride);
      ++g;
    }
  } else if (warp < 20) {
    // ================= epilogue: fp32 sum of the K-segments, then bias / activation / store =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GT_REGS_EPI));
    const int e = (warp - 4) >> 2;             // this warpgroup owns columns [64 e, 64 e + 64) of the tile
    const int r = tid & 127;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) &&
                     (EPI != GT_EPI_DACT || H == nullptr || (reinterpret_cast<uintptr_t>(H) & 15u) == 0);
    const int nseg = (nkc + seg - 1) / seg;
    int sg = 0;
    for (long long item = first; item < nitems; item += stride) {
      const int n_tile = (int)(item % nt);
      const int np = gt_np(N, n_tile);
      const int n0 = n_tile * GT_NMAX;
      const long long row = (item / nt) * GT_M + r;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      for (int q = 0; q < nseg; ++q, ++sg) {
        const int db = sg & 1;
        mbar_wait_hint(&d_full[db], (uint32_t)((sg >> 1) & 1));
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 64; c += 8) {       // 8 columns at a time: 64 running sums already fill the registers
          if (64 * e + c < np) {                // (16 at a time spilled the sums to local memory: profiles/r1_h)
            uint32_t u[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                         : "r"(lane_base + (uint32_t)db * GT_NMAX + 64 * e + c)
                         : "memory");
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[c + i] += __uint_as_float(u[i]);
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&d_free[db]);
      }
      if (row < M) {
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          const int col = n0 + 64 * e + c;
          if (64 * e + c >= np) continue;
          float* crow = C + row * ldc + col;
          if (vec && col + 16 <= N) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float o[4] = {acc[c + 4 * q4], acc[c + 4 * q4 + 1], acc[c + 4 * q4 + 2], acc[c + 4 * q4 + 3]};
              if (EPI == GT_EPI_BIAS_ACT) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col) + q4);
                o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
                if (apply_act) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) o[i] = act_forward(o[i], act);
                }
              } else if (apply_act) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(H + row * ldc + col) + q4);
                o[0] *= act_grad_from_output(h.x, act); o[1] *= act_grad_from_output(h.y, act);
                o[2] *= act_grad_from_output(h.z, act); o[3] *= act_grad_from_output(h.w, act);
              }
              reinterpret_cast<float4*>(crow)[q4] = make_float4(o[0], o[1], o[2], o[3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (col + i < N) {
                float o = acc[c + i];
                if (EPI == GT_EPI_BIAS_ACT) {
                  o += __ldg(bias + col + i);
                  if (apply_act) o = act_forward(o, act);
                } else if (apply_act) {
                  o *= act_grad_from_output(__ldg(H + row * ldc + col + i), act);
                }
                crow[i] = o;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512u);
}

}  // namespace molann
