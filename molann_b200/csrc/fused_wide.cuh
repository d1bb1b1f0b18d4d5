// fused_wide.cuh -- ONE persistent kernel for big systems with a wide first layer (C3 / C5: thousands of atoms per
// frame, d = 800 ... 2000 features -> [d, N1 <= 256, N2 <= 256, k <= 8]):  align -> features -> MLP per 128-frame tile,
// x read from HBM exactly once, features and activations never written to HBM, one launch per call.
//   reference: molann/ann.py:553-565 (PreprocessingANN.forward) -> :620-624 (MolANN.forward), MLP of :37-67.
//
// The layered path it replaces (general_forward, molann_b200.cu) ran preprocess -> [L, d] features in HBM -> pack +
// GEMM -> HBM -> pack + GEMM -> narrow kernel: six launches per 32768-frame chunk, 1.32x the algorithmic DRAM traffic
// (profiles/r2a_dram_launches_C3.csv) and no overlap between the HBM-bound geometry (185 us per chunk) and the
// tensor-core GEMMs (105 + 47 us).
//
// Problem: the frames arrive frame-major (a 24 KB frame carries ALL its 800 features) but the tensor core contracts
// K-chunk-major (a 16-wide K-chunk of ALL 128 frames of the tile), and a tile's features (128 x 3.2 KB) fit neither
// in shared memory nor in TMEM.  Two ideas make it one kernel anyway:
//   1. DEFERRED ROTATION.  Position features are (x_a - c) R with (c, R) known only after the whole frame's moments.
//      The geometry role therefore stores the RAW pivot-relative coordinates of the position atoms plus the frame's
//      moments (H, c); the rotation is solved later thread-per-frame (128 frames at once, no 1700-cycle serial chain
//      on the per-frame path) and applied when the operand chunk is built.  Bond / angle / dihedral features are
//      rigid-motion invariant and final as soon as they are computed.
//   2. A PER-CTA TRANSPOSITION SCRATCH IN L2.  Each CTA owns a small ring of 32-frame sub-tiles in global memory
//      (C3: 5 x 104 KB per CTA, 77 MB for the chip -- it is overwritten every tile, so it lives in the 126 MB L2 and
//      DRAM never sees it; x streams through with no reuse).  Geometry warps write frame rows, converter warps read
//      K-chunks.  Layout per sub-tile: [32 rows x 16 header floats][chunk][32 rows][16 floats].
// Internal feature order ("units" of 4 floats = 16 bytes, the K-major granule of the tensor-core operand): unit u <
// n_pos = (x, y, z of position atom u, invariant feature u); further invariant features follow four per unit.  With
// as many invariant columns as position atoms (C3, C5) K stays exactly d.  The first layer's weights are permuted,
// pre-scaled by the activation's exponent scale, split into TF32 hi / lo and laid out per K-chunk ONCE per plan
// (molann_b200_prepare), so a weight block is one bulk copy and no pack kernel runs in the steady state.
//
// Roles (28 warps, persistent CTA per SM, everything hands over through mbarriers):
//   X producer (1 warp)   cp.async.bulk of whole frames into a shared-memory ring
//   geometry   (4 warps)  warp = frame: pivoted moments (warp-shuffle sums), raw position atoms, invariant features
//                         -> scratch sub-tile;  no rotation here
//   converter  (4 warps)  thread = frame: quaternion rotation from the moments (polynomial fast path / Jacobi
//                         fallback, geometry.cuh) once per tile, then per K-chunk: 64 bytes from the scratch ->
//                         (p - c) R -> TF32 hi / lo (round to nearest) -> canonical K-major operand tile in smem
//   W producer (1 warp)   cp.async.bulk of the pre-packed weight block of each chunk (layer 1, then layer 2)
//   MMA        (1 warp)   elected lane: tcgen05.mma 3xTF32, SS form, M = 128, N = N1 (layer 1) / N2 (layer 2), K = 8;
//                         32-wide K segments into alternating TMEM accumulators (the tensor core truncates its fp32
//                         accumulator at every step: long sums in one accumulator cost 1e-5, gemm_tc.cuh)
//   epilogue   (16 warps) thread = row x 64 columns: fp32 sum of the segments in registers, bias + activation, the
//                         activations go straight back into the operand ring as layer 2's A chunks (hi / lo), second
//                         round of sums, activation, last (narrow) layer as register dot products, y.
#pragma once
#include "common.cuh"
#include "fused_tc.cuh"
#include "fused_ws.cuh"
#include "geometry.cuh"
#include "tc.cuh"

namespace molann {

constexpr int FW_M = 128;                 // frames per tile (tcgen05 M)
constexpr int FW_SUB = 32;                // frames per scratch sub-tile = rows of one converter warp
constexpr int FW_KU = 4;                  // 16-byte units per K-chunk
constexpr int FW_KC = 4 * FW_KU;          // K per chunk
constexpr int FW_SEGC = 2;                // chunks per accumulation segment (32 K = 12 MMA steps)
constexpr int FW_NMAX = 256;              // widest tensor-core layer
constexpr int FW_CW = 64;                 // accumulator columns per epilogue thread
constexpr int FW_WARPS = 28;
constexpr int FW_THREADS = FW_WARPS * 32;
constexpr int FW_W_EPI = 4, FW_W_WPROD = 20, FW_W_MMA = 21, FW_W_XPROD = 22, FW_W_GEO = 24;
constexpr int FW_NGW = 4;                 // geometry warps
// setmaxnreg budgets; pool = 896 threads x 72 registers = 64512 = 32 x (4 x 72 + 16 x 88 + 4 x 24 + 4 x 56)
constexpr int FW_REGS_CONV = 72, FW_REGS_EPI = 88, FW_REGS_CTRL = 24, FW_REGS_GEO = 56;
constexpr int FW_A_HALF = FW_M * FW_KC * 4;                               // one of hi / lo: 8 KB
constexpr int FW_STAGE_BYTES = 2 * FW_A_HALF + 2 * FW_NMAX * FW_KC * 4;   // 16 KB A + 32 KB W
constexpr int FW_MAX_STAGES = 4, FW_MAX_RING = 8, FW_MAX_SLOTS = 8;
constexpr int FW_HDR_FLOATS = FW_SUB * 16;                                // per sub-tile: 32 rows x (H[9], c[3], pad)
constexpr int FW_CHUNK_FLOATS = FW_SUB * FW_KC;                           // per sub-tile and chunk: 32 rows x 16

struct FwBars {
  unsigned long long empty[FW_MAX_STAGES], a_full[FW_MAX_STAGES], b_full[FW_MAX_STAGES];
  unsigned long long d_full[2], d_free[2];
  unsigned long long x_full[FW_MAX_RING], x_empty[FW_MAX_RING];
  unsigned long long s_full[FW_MAX_SLOTS], s_free[FW_MAX_SLOTS];
  unsigned long long turn[4];
  uint32_t tptr;
};

struct FwParams {
  // geometry tables in kernel order (device memory, inside the prepared buffer)
  const int* pos_atom;       // [n_pos]            local atom index of position unit u
  const int* align_idx;      // [n_align]
  const float* ref_x;        // [3 n_align]        centred reference
  const int* inv_ent;        // [n_inv_ent][6]     {type, a0, a1, a2, a3, first invariant column}
  int n_inp, n_align, n_pos, n_inv_ent, n_inv, n_units, use_angle;
  // MLP
  int n_hidden;              // tensor-core layers: 1 or 2
  int nkc1, n1p;             // layer 1: K-chunks, padded width
  int nkc2, n2p;             // layer 2 (n_hidden == 2): K-chunks (= n1p / 16), padded width
  int nlastp;                // padded width of the last hidden layer (row stride of w3)
  int kout;
  const float* w1p;          // [nkc1][2][KC/4][n1p][4]   hi block, lo block per chunk
  const float* w2p;          // [nkc2][2][KC/4][n2p][4]
  const float* b1s;          // [n1p] bias x activation scale
  const float* b2s;          // [n2p]
  const float* w3;           // [kout][nlastp]
  const float* b3;           // [kout]
  // transposition scratch (global memory, L2 resident): per CTA n_slots sub-tiles of slot_floats
  float* scratch;
  long long cta_floats;
  int slot_floats, n_slots;
  // shared memory
  int n_stages, n_ring, ring_slot_bytes;
  int off_stage, off_ring, off_b1, off_b2, off_w3, off_ypart;
  int off_pos, off_aidx, off_ref, off_ent;      // plan tables staged in smem (-1: read from global memory)
  int total_smem;
};

// activation on the pre-scaled pre-activation (weights and biases carry the exponent scale)
template <int ACT>
__device__ __forceinline__ float fw_act(float zs) { return ws_act<ACT>(zs); }

// writer of an invariant feature column into this frame's scratch row
struct FwInvOut {
  float* units;              // sub-tile's unit area
  int rr, n_pos;
  __device__ __forceinline__ void operator()(int v, float val) {
    const int idx = v < n_pos ? 4 * v + 3 : 3 * n_pos + v;          // unit * 4 + component
    const int u = idx >> 2;
    units[((u >> 2) * FW_SUB + rr) * FW_KC + (u & 3) * 4 + (idx & 3)] = val;
  }
};

__device__ __forceinline__ void fw_stage_step(int& s, uint32_t& par, int n_stages) {
  if (++s == n_stages) {
    s = 0;
    par ^= 1u;
  }
}
// stage index and use parity of global chunk number g
__device__ __forceinline__ void fw_stage_of(unsigned g, int n_stages, int& s, uint32_t& par) {
  const unsigned q = g / (unsigned)n_stages;
  s = (int)(g - q * (unsigned)n_stages);
  par = q & 1u;
}

// one row's 16 values of a K-chunk -> TF32 hi / lo -> canonical K-major operand tile (unit j at j * 2048 + row * 16)
__device__ __forceinline__ void fw_store_units(unsigned char* a_hi, int row, const float4 (&v)[FW_KU]) {
  unsigned char* a_lo = a_hi + FW_A_HALF;
#pragma unroll
  for (int j = 0; j < FW_KU; ++j) {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_tf32_rn(v[j].x, h0, l0);
    split_tf32_rn(v[j].y, h1, l1);
    split_tf32_rn(v[j].z, h2, l2);
    split_tf32_rn(v[j].w, h3, l3);
    *reinterpret_cast<uint4*>(a_hi + j * (FW_M * 16) + row * 16) = make_uint4(h0, h1, h2, h3);
    *reinterpret_cast<uint4*>(a_lo + j * (FW_M * 16) + row * 16) = make_uint4(l0, l1, l2, l3);
  }
}

// fp32 sum of `nseg` accumulator segments into acc[64] (columns col0 .. col0 + 63 of this thread's row)
__device__ __forceinline__ void fw_sum_segments(float (&acc)[FW_CW], int nseg, int col0, int np, uint32_t lane_base,
                                                FwBars* bars, unsigned& sg) {
#pragma unroll
  for (int i = 0; i < FW_CW; ++i) acc[i] = 0.f;
  for (int q = 0; q < nseg; ++q, ++sg) {
    const int db = (int)(sg & 1u);
    mbar_wait_hint(&bars->d_full[db], (sg >> 1) & 1u);
    tc_fence_after_sync();
#pragma unroll
    for (int c = 0; c < FW_CW; c += 8) {
      if (col0 + c < np) {
        uint32_t u[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(lane_base + (uint32_t)db * FW_NMAX + (uint32_t)(col0 + c))
                     : "memory");
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c + i] += __uint_as_float(u[i]);
      }
    }
    tc_fence_before_sync();
    mbar_arrive(&bars->d_free[db]);
  }
}

template <int ACT>
__global__ void __launch_bounds__(FW_THREADS, 1)
fused_wide_forward_kernel(const __grid_constant__ FwParams P, const float* __restrict__ x, float* __restrict__ y,
                          long long L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  FwBars* bars = reinterpret_cast<FwBars*>(smem);
  if (tid == 0) {
    for (int s = 0; s < FW_MAX_STAGES; ++s) {
      mbar_init(&bars->empty[s], 1);
      mbar_init(&bars->a_full[s], 128);
      mbar_init(&bars->b_full[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->d_full[s], 1);
      mbar_init(&bars->d_free[s], 16 * 32);
    }
    for (int s = 0; s < FW_MAX_RING; ++s) {
      mbar_init(&bars->x_full[s], 1);
      mbar_init(&bars->x_empty[s], 1);
    }
    for (int s = 0; s < FW_MAX_SLOTS; ++s) {
      mbar_init(&bars->s_full[s], FW_SUB);
      mbar_init(&bars->s_free[s], 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&bars->turn[s], 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&bars->tptr, 512u);
  // constants every role reads per tile: biases, last layer; plan tables when they fit
  {
    float* b1 = reinterpret_cast<float*>(smem + P.off_b1);
    for (int i = tid; i < P.n1p; i += FW_THREADS) b1[i] = __ldg(P.b1s + i);
    if (P.n_hidden == 2) {
      float* b2 = reinterpret_cast<float*>(smem + P.off_b2);
      for (int i = tid; i < P.n2p; i += FW_THREADS) b2[i] = __ldg(P.b2s + i);
    }
    float* w3 = reinterpret_cast<float*>(smem + P.off_w3);
    for (int i = tid; i < P.kout * P.nlastp; i += FW_THREADS) w3[i] = __ldg(P.w3 + i);
    if (P.off_pos >= 0) {
      int* t = reinterpret_cast<int*>(smem + P.off_pos);
      for (int i = tid; i < P.n_pos; i += FW_THREADS) t[i] = __ldg(P.pos_atom + i);
    }
    if (P.off_aidx >= 0) {
      int* t = reinterpret_cast<int*>(smem + P.off_aidx);
      for (int i = tid; i < P.n_align; i += FW_THREADS) t[i] = __ldg(P.align_idx + i);
    }
    if (P.off_ref >= 0) {
      float* t = reinterpret_cast<float*>(smem + P.off_ref);
      for (int i = tid; i < 3 * P.n_align; i += FW_THREADS) t[i] = __ldg(P.ref_x + i);
    }
    if (P.off_ent >= 0) {
      int* t = reinterpret_cast<int*>(smem + P.off_ent);
      for (int i = tid; i < ENTRY_INTS * P.n_inv_ent; i += FW_THREADS) t[i] = __ldg(P.inv_ent + i);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (bars->tptr != 0u) __trap();             // full allocation: base 0 (keeps tcgen05 addresses warp-uniform)

  const long long ntiles = (L + FW_M - 1) / FW_M;
  const int n3 = 3 * P.n_inp;
  const int nchunks_tile = P.nkc1 + (P.n_hidden == 2 ? P.nkc2 : 0);
  const int nseg1 = (P.nkc1 + FW_SEGC - 1) / FW_SEGC;
  const int nseg2 = P.n_hidden == 2 ? (P.nkc2 + FW_SEGC - 1) / FW_SEGC : 0;
  float* const cta_scratch = P.scratch + (long long)blockIdx.x * P.cta_floats;
  unsigned char* const stages = smem + P.off_stage;
  unsigned char* const ring = smem + P.off_ring;

  if (warp >= FW_W_GEO) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_GEO));
  else if (warp >= FW_W_WPROD) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_CTRL));

  if (warp == FW_W_XPROD) {
    // ================= X producer: whole frames into the shared-memory ring =================
    int rs = 0;
    uint32_t rpar = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int r = 0; r < FW_M; ++r) {
        const long long f = tile * FW_M + r;
        if (f >= L) break;
        const float* src = x + f * n3;
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint32_t bytes = ((uint32_t)n3 * 4u + off + 15u) & ~15u;
        unsigned char* dst = ring + (size_t)rs * P.ring_slot_bytes;
        if (lane == 0) mbar_wait_hint(&bars->x_empty[rs], rpar ^ 1u);
        __syncwarp();
        // the bulk copy starts at the 16-byte boundary below the frame and may run up to 15 bytes past it:
        // never past the end of x (last frame of the batch unless it ends on the grid)
        if (f + 1 < L || ((off + (uint32_t)n3 * 4u) & 15u) == 0u) {
          if (lane == 0) {
            mbar_expect_tx(&bars->x_full[rs], bytes);
            bulk_g2s(dst, reinterpret_cast<const unsigned char*>(src) - off, bytes, &bars->x_full[rs]);
          }
        } else {
          float* d = reinterpret_cast<float*>(dst + off);
          for (int i = lane; i < n3; i += 32) d[i] = __ldg(src + i);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->x_full[rs]);
        }
        if (++rs == P.n_ring) {
          rs = 0;
          rpar ^= 1u;
        }
      }
    }
  } else if (warp >= FW_W_GEO) {
    // ================= geometry: moments, raw position atoms, invariant features -> scratch =================
    const int gw = warp - FW_W_GEO;
    const int* pos_atom = P.off_pos >= 0 ? reinterpret_cast<const int*>(smem + P.off_pos) : P.pos_atom;
    const int* aidx = P.off_aidx >= 0 ? reinterpret_cast<const int*>(smem + P.off_aidx) : P.align_idx;
    const float* refx = P.off_ref >= 0 ? reinterpret_cast<const float*>(smem + P.off_ref) : P.ref_x;
    const int* ient = P.off_ent >= 0 ? reinterpret_cast<const int*>(smem + P.off_ent) : P.inv_ent;
    const bool aligned = P.n_align > 0;
    const float inv_n = aligned ? 1.0f / (float)P.n_align : 0.f;
    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      for (int r = gw; r < FW_M; r += FW_NGW) {
        const long long f = tile * FW_M + r;
        // scratch sub-tile of this frame
        const unsigned j = it * (FW_M / FW_SUB) + (unsigned)(r / FW_SUB);
        const unsigned use = j / (unsigned)P.n_slots;
        const int slot = (int)(j - use * (unsigned)P.n_slots);
        const int rr = r & (FW_SUB - 1);
        if (rr < FW_NGW) {                           // this warp's first frame in the sub-tile: the slot must be free
          if (lane == 0) mbar_wait_hint(&bars->s_free[slot], (use & 1u) ^ 1u);
          __syncwarp();
        }
        if (f < L) {
          float* sub = cta_scratch + (long long)slot * P.slot_floats;
          float* units = sub + FW_HDR_FLOATS;
          // frame in the ring
          const unsigned ic = it * FW_M + (unsigned)r;
          const unsigned ruse = ic / (unsigned)P.n_ring;
          const int rs = (int)(ic - ruse * (unsigned)P.n_ring);
          if (lane == 0) mbar_wait_hint(&bars->x_full[rs], ruse & 1u);
          __syncwarp();
          const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(x + f * n3) & 15u);
          const float* xf = reinterpret_cast<const float*>(ring + (size_t)rs * P.ring_slot_bytes + off);
          float pvx = 0.f, pvy = 0.f, pvz = 0.f;
          if (aligned) {
            // pivoted one-pass moments (reference ann.py:179-187): d_k = x_k - x_{A_0}, H = sum d_k^T y_k (the
            // reference is centred), c_rel = mean d_k
            const float* p0 = xf + 3 * aidx[0];
            pvx = p0[0]; pvy = p0[1]; pvz = p0[2];
            float m[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) m[i] = 0.f;
            for (int k = lane; k < P.n_align; k += 32) {
              const float* p = xf + 3 * aidx[k];
              const float px = p[0] - pvx, py = p[1] - pvy, pz = p[2] - pvz;
              const float y0 = refx[3 * k], y1 = refx[3 * k + 1], y2 = refx[3 * k + 2];
              m[0] = fmaf(px, y0, m[0]); m[1] = fmaf(px, y1, m[1]); m[2] = fmaf(px, y2, m[2]);
              m[3] = fmaf(py, y0, m[3]); m[4] = fmaf(py, y1, m[4]); m[5] = fmaf(py, y2, m[5]);
              m[6] = fmaf(pz, y0, m[6]); m[7] = fmaf(pz, y1, m[7]); m[8] = fmaf(pz, y2, m[8]);
              m[9] += px; m[10] += py; m[11] += pz;
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) m[i] = gsum<32>(m[i]);
            if (lane == 0) {
              float4* h = reinterpret_cast<float4*>(sub + rr * 16);
              h[0] = make_float4(m[0], m[1], m[2], m[3]);
              h[1] = make_float4(m[4], m[5], m[6], m[7]);
              h[2] = make_float4(m[8], m[9] * inv_n, m[10] * inv_n, m[11] * inv_n);
            }
          }
          // raw (pivot-relative) coordinates of the position atoms; the w slot of a unit belongs to invariant
          // column u when there is one, else it is zero
          for (int u = lane; u < P.n_pos; u += 32) {
            const float* p = xf + 3 * pos_atom[u];
            float* dst = units + ((u >> 2) * FW_SUB + rr) * FW_KC + (u & 3) * 4;
            *reinterpret_cast<float2*>(dst) = make_float2(p[0] - pvx, p[1] - pvy);
            dst[2] = p[2] - pvz;
            if (u >= P.n_inv) dst[3] = 0.f;
          }
          // invariant features (bond / angle / dihedral, ann.py:323-351) from the raw coordinates
          FwInvOut out{units, rr, P.n_pos};
          Rigid none;
          for (int e = lane; e < P.n_inv_ent; e += 32) {
            const Entry en = load_entry(ient + ENTRY_INTS * e);
            feature_forward(en, xf, false, none, P.use_angle, out);
          }
          if (lane == 0 && P.n_inv > P.n_pos) {       // zero the unused tail of the last invariant unit
            for (int v = P.n_inv; ((v - P.n_pos) & 3) != 0; ++v) out(v, 0.f);
          }
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bars->x_empty[rs]);
            mbar_arrive(&bars->s_full[slot]);
          }
        } else if (lane == 0) {
          mbar_arrive(&bars->s_full[slot]);           // past the end of the batch: keep the sub-tile's count whole
        }
      }
    }
  } else if (warp < FW_W_EPI) {
    // ================= converter: rotation per frame, then K-chunks scratch -> TF32 hi / lo operand tiles ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FW_REGS_CONV));
    const int row = tid;                               // 0 .. 127
    const bool aligned = P.n_align > 0;
    unsigned it = 0, g = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      // layer 2's operand chunks of the previous tile share the stage ring: they must all have been written before
      // this role asks for a stage again (a waiter may be at most one phase ahead of an mbarrier)
      if (it > 0 && P.n_hidden == 2) mbar_wait_hint(&bars->turn[3], (it - 1u) & 1u);
      const unsigned j = it * (FW_M / FW_SUB) + (unsigned)warp;
      const unsigned use = j / (unsigned)P.n_slots;
      const int slot = (int)(j - use * (unsigned)P.n_slots);
      mbar_wait_hint(&bars->s_full[slot], use & 1u);
      const float* sub = cta_scratch + (long long)slot * P.slot_floats;
      const float4* units = reinterpret_cast<const float4*>(sub + FW_HDR_FLOATS) + lane * FW_KU;
      const bool valid = tile * FW_M + row < L;
      Rigid rg;
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (aligned) {
        const float4* h = reinterpret_cast<const float4*>(sub + lane * 16);
        float4 h0 = make_float4(1.f, 0.f, 0.f, 0.f), h1 = make_float4(1.f, 0.f, 0.f, 0.f),
               h2 = make_float4(1.f, 0.f, 0.f, 0.f);
        if (valid) { h0 = __ldcg(h); h1 = __ldcg(h + 1); h2 = __ldcg(h + 2); }
        rg.H[0] = h0.x; rg.H[1] = h0.y; rg.H[2] = h0.z; rg.H[3] = h0.w;
        rg.H[4] = h1.x; rg.H[5] = h1.y; rg.H[6] = h1.z; rg.H[7] = h1.w; rg.H[8] = h2.x;
        c0 = h2.y; c1 = h2.z; c2 = h2.w;
        kabsch_rotation(rg);                          // reference ann.py:188-195 as a quaternion eigenproblem
      }
      float4 nxt[FW_KU];
#pragma unroll
      for (int q = 0; q < FW_KU; ++q) nxt[q] = valid ? __ldcg(units + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kc = 0; kc < P.nkc1; ++kc, ++g) {
        float4 v[FW_KU];
#pragma unroll
        for (int q = 0; q < FW_KU; ++q) v[q] = nxt[q];
        if (kc + 1 < P.nkc1 && valid) {
          const float4* src = units + (size_t)(kc + 1) * (FW_CHUNK_FLOATS / 4);
#pragma unroll
          for (int q = 0; q < FW_KU; ++q) nxt[q] = __ldcg(src + q);
        }
#pragma unroll
        for (int q = 0; q < FW_KU; ++q) {
          const int u = kc * FW_KU + q;
          if (u >= P.n_units) {
            v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (aligned && u < P.n_pos) {        // z = (p - c) R  (ann.py:197), p and c relative to the pivot
            const float dx = v[q].x - c0, dy = v[q].y - c1, dz = v[q].z - c2;
            v[q].x = fmaf(dx, rg.R[0], fmaf(dy, rg.R[3], dz * rg.R[6]));
            v[q].y = fmaf(dx, rg.R[1], fmaf(dy, rg.R[4], dz * rg.R[7]));
            v[q].z = fmaf(dx, rg.R[2], fmaf(dy, rg.R[5], dz * rg.R[8]));
          }
        }
        int s;
        uint32_t par;
        fw_stage_of(g, P.n_stages, s, par);
        mbar_wait_hint(&bars->empty[s], par ^ 1u);
        fw_store_units(stages + (size_t)s * FW_STAGE_BYTES, row, v);
        fence_proxy_async_smem();
        mbar_arrive(&bars->a_full[s]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[slot]);
      g += (unsigned)(nchunks_tile - P.nkc1);
    }
  } else if (warp == FW_W_WPROD) {
    // ================= W producer: pre-packed weight blocks, layer 1 then layer 2 =================
    if (lane == 0) {
      const uint32_t bytes1 = 2u * FW_KC * (uint32_t)P.n1p * 4u, bytes2 = 2u * FW_KC * (uint32_t)P.n2p * 4u;
      int s = 0;
      uint32_t par = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks_tile; ++kc) {
          const bool l1 = kc < P.nkc1;
          const float* src = l1 ? P.w1p + (size_t)kc * (2 * FW_KC * P.n1p)
                                : P.w2p + (size_t)(kc - P.nkc1) * (2 * FW_KC * P.n2p);
          const uint32_t bytes = l1 ? bytes1 : bytes2;
          mbar_wait_hint(&bars->empty[s], par ^ 1u);
          mbar_expect_tx(&bars->b_full[s], bytes);
          bulk_g2s(stages + (size_t)s * FW_STAGE_BYTES + 2 * FW_A_HALF, src, bytes, &bars->b_full[s]);
          fw_stage_step(s, par, P.n_stages);
        }
      }
    }
  } else if (warp == FW_W_MMA) {
    // ================= MMA issuer =================
    const uint32_t leader = elect_one();
    int s = 0;
    uint32_t par = 0;
    unsigned sg = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int layer = 0; layer < P.n_hidden; ++layer) {
        const int nkc = layer == 0 ? P.nkc1 : P.nkc2;
        const int np = layer == 0 ? P.n1p : P.n2p;
        const uint32_t idesc = idesc_tf32(FW_M, np);
        const uint32_t lbo_b = (uint32_t)np * 16u;
        for (int kc = 0; kc < nkc; ++kc) {
          const int db = (int)(sg & 1u);
          const uint32_t d = (uint32_t)db * FW_NMAX;
          const bool seg_first = (kc % FW_SEGC) == 0;
          const bool seg_last = (kc % FW_SEGC) == FW_SEGC - 1 || kc == nkc - 1;
          if (seg_first) {
            mbar_wait_hint(&bars->d_free[db], ((sg >> 1) & 1u) ^ 1u);
            tc_fence_after_sync();
          }
          mbar_wait_hint(&bars->a_full[s], par);
          mbar_wait_hint(&bars->b_full[s], par);
          tc_fence_after_sync();
          const uint32_t a_hi = smem_u32(stages + (size_t)s * FW_STAGE_BYTES), a_lo = a_hi + FW_A_HALF;
          const uint32_t b_hi = a_hi + 2 * FW_A_HALF, b_lo = b_hi + (uint32_t)np * FW_KC * 4u;
          // cross terms first, leading terms last: in a fresh accumulator only the leading steps round at full scale
#pragma unroll 1
          for (int jj = 0; jj < FW_KC / 8; ++jj) {
            const uint64_t ah = smem_desc_kmajor(a_hi + jj * (2u * FW_M * 16u), FW_M * 16u, 128);
            const uint64_t al = smem_desc_kmajor(a_lo + jj * (2u * FW_M * 16u), FW_M * 16u, 128);
            const uint64_t bh = smem_desc_kmajor(b_hi + jj * (2u * lbo_b), lbo_b, 128);
            const uint64_t bl = smem_desc_kmajor(b_lo + jj * (2u * lbo_b), lbo_b, 128);
            if (leader) {
              mma_tf32_ss(d, al, bh, idesc, (!seg_first || jj > 0) ? 1u : 0u);
              mma_tf32_ss(d, ah, bl, idesc, 1);
            }
          }
#pragma unroll 1
          for (int jj = 0; jj < FW_KC / 8; ++jj) {
            const uint64_t ah = smem_desc_kmajor(a_hi + jj * (2u * FW_M * 16u), FW_M * 16u, 128);
            const uint64_t bh = smem_desc_kmajor(b_hi + jj * (2u * lbo_b), lbo_b, 128);
            if (leader) mma_tf32_ss(d, ah, bh, idesc, 1);
          }
          if (leader) mma_commit(&bars->empty[s]);
          if (seg_last) {
            if (leader) mma_commit(&bars->d_full[db]);
            ++sg;
          }
          __syncwarp();
          fw_stage_step(s, par, P.n_stages);
        }
      }
    }
  } else if (warp < FW_W_WPROD) {
    // ================= epilogue: segment sums, activations, layer 2 operand chunks, last layer, y =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FW_REGS_EPI));
    const int e = (warp - FW_W_EPI) >> 2;            // this warpgroup owns accumulator columns [64 e, 64 e + 64)
    const int row = tid & 127;
    const int col0 = FW_CW * e;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    const float* b1 = reinterpret_cast<const float*>(smem + P.off_b1);
    const float* b2 = reinterpret_cast<const float*>(smem + P.off_b2);
    const float* w3 = reinterpret_cast<const float*>(smem + P.off_w3);
    float* ypart = reinterpret_cast<float*>(smem + P.off_ypart);
    unsigned sg = 0, it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      float acc[FW_CW];
      fw_sum_segments(acc, nseg1, col0, P.n1p, lane_base, bars, sg);
#pragma unroll
      for (int c = 0; c < FW_CW; ++c)
        if (col0 + c < P.n1p) acc[c] = fw_act<ACT>(acc[c] + b1[col0 + c]);
      int nlast = P.n1p;
      if (P.n_hidden == 2) {
        // h1 goes back into the operand ring as layer 2's A chunks, warpgroup after warpgroup (chunk order)
        if (e > 0) mbar_wait_hint(&bars->turn[e - 1], it & 1u);
        const unsigned gbase = it * (unsigned)nchunks_tile + (unsigned)P.nkc1;
#pragma unroll
        for (int q = 0; q < FW_CW / FW_KC; ++q) {
          const int kc2 = (FW_CW / FW_KC) * e + q;
          if (kc2 < P.nkc2) {
            int s;
            uint32_t par;
            fw_stage_of(gbase + (unsigned)kc2, P.n_stages, s, par);
            mbar_wait_hint(&bars->empty[s], par ^ 1u);
            float4 v[FW_KU];
#pragma unroll
            for (int jx = 0; jx < FW_KU; ++jx)
              v[jx] = make_float4(acc[FW_KC * q + 4 * jx], acc[FW_KC * q + 4 * jx + 1], acc[FW_KC * q + 4 * jx + 2],
                                  acc[FW_KC * q + 4 * jx + 3]);
            fw_store_units(stages + (size_t)s * FW_STAGE_BYTES, row, v);
            fence_proxy_async_smem();
            mbar_arrive(&bars->a_full[s]);
          }
        }
        mbar_arrive(&bars->turn[e]);
        fw_sum_segments(acc, nseg2, col0, P.n2p, lane_base, bars, sg);
#pragma unroll
        for (int c = 0; c < FW_CW; ++c)
          if (col0 + c < P.n2p) acc[c] = fw_act<ACT>(acc[c] + b2[col0 + c]);
        nlast = P.n2p;
      }
      // last (narrow) layer: partial dot products over this thread's columns, summed in a fixed order
      for (int o = 0; o < P.kout; ++o) {
        float pd = 0.f;
#pragma unroll
        for (int c = 0; c < FW_CW; ++c)
          if (col0 + c < nlast) pd = fmaf(acc[c], w3[o * P.nlastp + col0 + c], pd);
        ypart[(e * FW_M + row) * P.kout + o] = pd;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(16 * 32) : "memory");
      if (e == 0) {
        const long long f = tile * FW_M + row;
        if (f < L) {
          for (int o = 0; o < P.kout; ++o) {
            float v = __ldg(P.b3 + o);
#pragma unroll
            for (int ee = 0; ee < 4; ++ee) v += ypart[(ee * FW_M + row) * P.kout + o];
            y[f * P.kout + o] = v;
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(16 * 32) : "memory");
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512u);
}

// ---------------------------------------------------------------------------------------------------------------
// One-time packing (molann_b200_prepare): out[kc][hi | lo][(k / 4)][np][4] of  scale * W[n][colmap[k]]
// ---------------------------------------------------------------------------------------------------------------
__global__ void fw_pack_kernel(const float* __restrict__ W, int ldw, int N, const int* __restrict__ colmap, int Kp,
                               int np, float scale, float* __restrict__ out) {
  const long long total = (long long)np * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kp), k = (int)(i - (long long)n * Kp);
    const int col = colmap != nullptr ? __ldg(colmap + k) : (k < ldw ? k : -1);
    const float w = (n < N && col >= 0) ? scale * __ldg(W + (long long)n * ldw + col) : 0.f;
    uint32_t hi, lo;
    split_tf32_rn(w, hi, lo);
    lo = (lo + 0x1000u) & 0xffffe000u;
    const int kc = k / FW_KC, kk = k - kc * FW_KC;
    float* blk = out + (size_t)kc * (2 * FW_KC * np);
    const int off = ((kk >> 2) * np + n) * 4 + (kk & 3);
    blk[off] = __uint_as_float(hi);
    blk[FW_KC * np + off] = __uint_as_float(lo);
  }
}
// dst[i] = i < n ? scale * src[i] : 0   (biases), and the last layer's rows padded to `ldp`
__global__ void fw_pack_small_kernel(const float* __restrict__ b1, int n1, int n1p, float s1, float* __restrict__ o1,
                                     const float* __restrict__ b2, int n2, int n2p, float s2, float* __restrict__ o2,
                                     const float* __restrict__ w3, int kout, int nlast, int nlastp,
                                     float* __restrict__ o3, const float* __restrict__ b3, float* __restrict__ ob3) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = t; i < n1p; i += nt) o1[i] = i < n1 ? s1 * __ldg(b1 + i) : 0.f;
  if (b2 != nullptr)
    for (int i = t; i < n2p; i += nt) o2[i] = i < n2 ? s2 * __ldg(b2 + i) : 0.f;
  for (int i = t; i < kout * nlastp; i += nt) {
    const int o = i / nlastp, c = i - o * nlastp;
    o3[i] = c < nlast ? __ldg(w3 + o * nlast + c) : 0.f;
  }
  for (int i = t; i < kout; i += nt) ob3[i] = __ldg(b3 + i);
}

}  // namespace molann
