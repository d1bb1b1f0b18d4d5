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
//      (C3: 6 x 108 KB per CTA, 96 MB for the chip -- it is overwritten every tile, so most of it lives in the 126 MB
//      L2; x streams through with an evict-first hint, the scratch stores carry evict-last and the converter discards
//      a row's lines once it has read them).  The geometry warps write their results STRAIGHT into the frame's row
//      (small L2 stores); converter threads pull 64 bytes of their row per K-chunk through a cp.async staging ring.
//      Row = [moment partials of the frame's geometry warps][units], padded to whole 128-byte lines.
// Internal feature order ("units" of 4 floats = 16 bytes, the K-major granule of the tensor-core operand): unit u <
// n_pos = (x, y, z of position atom u, invariant feature u); further invariant features follow four per unit.  With
// as many invariant columns as position atoms (C3, C5) K stays exactly d.  The first layer's weights are permuted,
// pre-scaled by the activation's exponent scale, split into TF32 hi / lo and laid out per K-chunk ONCE per plan
// (molann_b200_prepare), so a weight block is one bulk copy and no pack kernel runs in the steady state.
//
// Roles (32 warps, persistent CTA per SM, everything hands over through mbarriers; DESIGN.md 3.6 has the measurements
// behind each choice):
//   X producers (2 warps) cp.async (16 B, L2 evict-first) of whole frames into a shared-memory ring, alternate 2 KB pieces
//   geometry   (8 warps)  two groups of four warps on alternate frames; the warps of a group share ONE frame (thread =
//                         alignment atom / position atom / invariant entry) and never meet: pivoted moments -> the
//                         warp's own 12-float partial in the row header, raw position atoms, invariant features ->
//                         the row; the ring slot is released as soon as the frame has been read;  no rotation here
//   converter  (4 warps)  thread = frame: sum of the moment partials, quaternion rotation (polynomial fast path /
//                         Jacobi fallback, geometry.cuh) once per tile, then per K-chunk: 64 bytes of the row ->
//                         (p - c) R -> TF32 hi / lo (round to nearest) -> canonical K-major operand tile in smem
//   W producer (1 warp)   cp.async.bulk of the pre-packed weight block of each chunk (layer 1, then layer 2)
//   MMA        (1 warp)   elected lane: tcgen05.mma 3xTF32, SS form, M = 128, N = N1 (layer 1) / N2 (layer 2), K = 8.
//                         Layer 1: 64-wide K segments into two alternating 256-column TMEM accumulators (the tensor
//                         core truncates its fp32 accumulator at every step: long sums in one accumulator cost 1e-5,
//                         gemm_tc.cuh).  Layer 2 (K <= 256): the chunks of one epilogue warpgroup go into one of
//                         four 128-column accumulators (two of 256 when N2 > 128) -- 24 steps each and no drain
//                         between its MMAs, so the epilogue can hand all of h1 over before it reads anything back
//   epilogue   (16 warps) thread = row x 64 columns: fp32 sum of the segments in registers, bias + activation, the
//                         activations go straight back into the operand ring as layer 2's A chunks (hi / lo), sum of
//                         the layer-2 accumulators, activation, last (narrow) layer as register dot products, y
//                         (STORE_H instantiation: also h1 / h2 for the value-and-gradient path).
// Template parameters: ACT, KU (16-byte units per K-chunk: 4, or 2 = half-size operand stages for 60 KB frames),
// STORE_H.  FW_OPT_* / FW_DBG_* are the A/B and timing-ablation switches of tests/cuda/fw_trace.cu (profiles/r3i, r3u).
#pragma once
#include "common.cuh"
#include "fused_tc.cuh"
#include "fused_ws.cuh"
#include "geometry.cuh"
#include "staged_block.cuh"
#include "tc.cuh"

// A/B switches of tests/cuda/fw_trace.cu (defaults = the shipped configuration)
#ifndef FW_OPT_POLICY
#define FW_OPT_POLICY 1      // evict-last scratch stores + discard of the rows the converter has read
#endif
#ifndef FW_OPT_XPROD2
#define FW_OPT_XPROD2 1      // two X producer warps, four copies per loop trip (0: one warp, rolled loop)
#endif
#ifndef FW_OPT_LATE
#define FW_OPT_LATE 1        // release the frame's ring slot before the dihedral arithmetic
#endif
#ifndef FW_SEG_K
#define FW_SEG_K 64          // K per accumulation segment of layer 1 (64 K = 24 MMA steps into a fresh accumulator, the
                             // same count as a layer-2 accumulator; 32 -> 64: +2 %, y error unchanged at 1e-6)
#endif
#ifndef FW_OPT_LD2
#define FW_OPT_LD2 0         // segment drain: two TMEM loads per wait
#endif
#ifndef FW_OPT_ACC2
#define FW_OPT_ACC2 0        // segment drain: packed f32x2 adds
#endif

namespace molann {

constexpr int FW_M = 128;                 // frames per tile (tcgen05 M)
constexpr int FW_SUB = 32;                // frames per scratch sub-tile = rows of one converter warp
// K-chunk size: template parameter KU of the kernel = 16-byte units per chunk (4: K = 16 per operand stage, the
// default; 2: K = 8, half-size stages for frames so large -- C5: 60 KB -- that two of them and two full stages do not
// fit in shared memory).  An accumulation segment is always 32 K = 12 MMA steps.
__host__ __device__ constexpr int fw_kc(int ku) { return 4 * ku; }
constexpr int FW_NMAX = 256;              // widest tensor-core layer
constexpr int FW_CW = 64;                 // accumulator columns per epilogue thread
constexpr int FW_WARPS = 32;
constexpr int FW_THREADS = FW_WARPS * 32;
constexpr int FW_W_EPI = 4, FW_W_WPROD = 20, FW_W_MMA = 21, FW_W_XPROD = 22, FW_W_GEO = 24;   // X producer: warps 22, 23
constexpr int FW_GEO_WARPS = 8;           // geometry warps: FW_NGG groups (alternate frames) of FW_NGW warps (one frame)
constexpr int FW_NGG = 2, FW_NGW = FW_GEO_WARPS / FW_NGG;
__host__ __device__ constexpr int fw_conv_chunk(int ku) { return FW_M * fw_kc(ku) * 4; }   // converter staging: one raw K-chunk, [unit][row][16 B]
constexpr int FW_MAX_CDEPTH = 4;
// setmaxnreg budgets; pool = 1024 threads x 64 registers = 65536 = 32 x (4 x 56 + 16 x 80 + 4 x 24 + 8 x 56).  The
// geometry role bounds the kernel (tests/cuda/fw_trace.cu), so it gets two groups of spill-free warps; the MMA chain
// has slack and pays for it with a few spill reloads per chunk.
constexpr int FW_REGS_CONV = 56, FW_REGS_EPI = 80, FW_REGS_CTRL = 24, FW_REGS_GEO = 56;
static_assert(4 * FW_REGS_CONV + 16 * FW_REGS_EPI + 4 * FW_REGS_CTRL + 8 * FW_REGS_GEO <= 32 * 64, "register pool");
constexpr uint32_t FW_PIECE = 2048;      // bytes per bulk copy of the frame ring
constexpr int FW_REGS_LAUNCH = 64;        // 65536 / 1024 threads: what the CTA starts with
__host__ __device__ constexpr int fw_a_half(int ku) { return FW_M * fw_kc(ku) * 4; }           // one of hi / lo: 8 KB (KU = 4)
__host__ __device__ constexpr int fw_stage_bytes(int ku) { return 2 * fw_a_half(ku) + 2 * FW_NMAX * fw_kc(ku) * 4; }   // 16 KB A + 32 KB W
constexpr int FW_MAX_STAGES = 4, FW_MAX_RING = 8, FW_MAX_SLOTS = 8;
// per scratch row: one 12-float moment partial (H[9], sum d[3]) per geometry warp of the frame's group; the converter
// adds them in a fixed order
constexpr int FW_HDR_FLOATS = 12 * FW_NGW;
// a row = header + nkc1 K-chunks, rounded up to whole 128-byte L2 lines (no line is shared by two rows)
__host__ __device__ inline int fw_row_floats(int nkc1, int ku) { return (FW_HDR_FLOATS + nkc1 * fw_kc(ku) + 31) & ~31; }

// Development aid (tests/cuda/fw_trace.cu): clock64() stamps of CTA 0, one lane per role.  Compiled out of the product.
#ifdef MOLANN_WS_TRACE
__device__ long long g_fw_trace[4 * 256 * 8];       // [role][item][event]
#define FW_EVT(role, i, ev)                                                                      \
  do {                                                                                           \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (i) < 256) g_fw_trace[((role) * 256 + (i)) * 8 + (ev)] = clock64(); \
  } while (0)
// per-tile steady-state record: [role][tile][0] = time stamp, [1..] = cumulative wait cycles of that role
__device__ long long g_fw_tiles[4 * 32 * 8];
#define FW_TILE(role, it, ev, val)                                                               \
  do {                                                                                           \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it) < 32) g_fw_tiles[((role) * 32 + (it)) * 8 + (ev)] = (val); \
  } while (0)
#define FW_WAIT_BEGIN() const long long fw_t0_ = clock64()
#define FW_WAIT_END(acc) (acc) += clock64() - fw_t0_
#define FW_TRACE_ONLY(x) x
#else
#define FW_EVT(role, i, ev) \
  do {                      \
  } while (0)
#define FW_TILE(role, it, ev, val) \
  do {                             \
  } while (0)
#define FW_WAIT_BEGIN() \
  do {                  \
  } while (0)
#define FW_WAIT_END(acc) \
  do {                   \
  } while (0)
#define FW_TRACE_ONLY(x)
#endif

struct FwBars {
  unsigned long long empty[FW_MAX_STAGES], a_full[FW_MAX_STAGES], b_full[FW_MAX_STAGES];
  unsigned long long d_full[2], d_free[2];
  unsigned long long x_full[FW_MAX_RING], x_empty[FW_MAX_RING];
  unsigned long long s_full[FW_MAX_SLOTS], s_free[FW_MAX_SLOTS];
  unsigned long long turn[4];
  unsigned long long l2_full, l2_free;
  uint32_t tptr;
};

struct FwParams {
  // geometry tables in kernel order (device memory, inside the prepared buffer)
  const int* pos_atom;       // [n_pos]            local atom index of position unit u
  const int* align_idx;      // [n_align]
  const float* ref_x;        // [3 n_align]        centred reference
  const int* inv_ent;        // [n_inv_ent][6]     {type, a0, a1, a2, a3, first invariant column}
  int n_inp, n_align, n_pos, n_inv_ent, n_inv, n_units, use_angle;
  int pos_is_align;          // the position atoms ARE the alignment selection, in the same order (one pass serves both)
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
  float* h1_out;             // optional [L][n1] / [L][n2]: the hidden activations, for the value-and-gradient path
  float* h2_out;
  int n1, n2;                // true (unpadded) widths = row strides of h1_out / h2_out
  const float* x_base;       // = x (the geometry role only needs its low address bits: where a frame sits in its ring slot)
  // transposition scratch (global memory, L2 resident): per CTA n_slots sub-tiles of slot_floats
  float* scratch;
  long long cta_floats;
  int slot_floats, n_slots, row_floats;
  // shared memory
  int n_stages, n_ring, ring_slot_bytes;
  int conv_depth;            // raw K-chunks the converter keeps in flight (cp.async staging ring)
  int off_stage, off_ring, off_b1, off_b2, off_w3, off_ypart, off_cstage;
  int off_pos, off_aidx, off_ref, off_ent;      // plan tables staged in smem (-1: read from global memory)
  int total_smem;
};

// activation on the pre-scaled pre-activation (weights and biases carry the exponent scale)
template <int ACT>
__device__ __forceinline__ float fw_act(float zs) { return ws_act<ACT>(zs); }

// Scratch stores carry an evict-last L2 policy: a row must stay in L2 until the converter has read it (48 % of the
// plain stores missed L2 and 88 % of the scratch was written back to DRAM while the evict-first frame stream passed
// through: profiles/r3g).  The converter discards the lines when it is done with them.
__device__ __forceinline__ void fw_st1(float* p, float v, unsigned long long pol) {
#if FW_OPT_POLICY
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
  __stcg(p, v);
#endif
}
__device__ __forceinline__ void fw_st2(float* p, float a, float b, unsigned long long pol) {
#if FW_OPT_POLICY
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
#else
  __stcg(reinterpret_cast<float2*>(p), make_float2(a, b));
#endif
}
// writer of an invariant feature column into this frame's scratch row (global memory, L2)
struct FwInvOut {
  float* units;              // unit area of the frame's scratch row
  int n_pos;
  unsigned long long pol;
  __device__ __forceinline__ void operator()(int v, float val) {
    fw_st1(units + (v < n_pos ? 4 * v + 3 : 3 * n_pos + v), val, pol);   // unit * 4 + component
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
// ... `by` chunks further on
__device__ __forceinline__ void fw_stage_skip(int& s, uint32_t& par, int by, int n_stages) {
  const unsigned t = (unsigned)s + (unsigned)by;
  const unsigned q = t / (unsigned)n_stages;
  s = (int)(t - q * (unsigned)n_stages);
  par ^= q & 1u;
}

// one row's 16 values of a K-chunk -> TF32 hi / lo -> canonical K-major operand tile (unit j at j * 2048 + row * 16)
template <int KU>
__device__ __forceinline__ void fw_store_units(unsigned char* a_hi, int row, const float4 (&v)[KU]) {
  unsigned char* a_lo = a_hi + fw_a_half(KU);
#pragma unroll
  for (int j = 0; j < KU; ++j) {
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
                                                FwBars* bars, unsigned& sg, bool trace) {
#pragma unroll
  for (int i = 0; i < FW_CW; ++i) acc[i] = 0.f;
  for (int q = 0; q < nseg; ++q, ++sg) {
    const int db = (int)(sg & 1u);
    if (trace) FW_EVT(2, 16 + q, 0);
    mbar_wait_hint(&bars->d_full[db], (sg >> 1) & 1u);
    tc_fence_after_sync();
    if (trace) FW_EVT(2, 16 + q, 1);
#if FW_OPT_LD2
#pragma unroll
    for (int c = 0; c < FW_CW; c += 16) {
      if (col0 + c < np) {
        uint32_t u[8], v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(lane_base + (uint32_t)db * FW_NMAX + (uint32_t)(col0 + c))
                     : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(lane_base + (uint32_t)db * FW_NMAX + (uint32_t)(col0 + c + 8))
                     : "memory");
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c + i] += __uint_as_float(u[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c + 8 + i] += __uint_as_float(v[i]);
      }
    }
#else
#pragma unroll
    for (int c = 0; c < FW_CW; c += 8) {
      if (col0 + c < np) {
        uint32_t u[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(lane_base + (uint32_t)db * FW_NMAX + (uint32_t)(col0 + c))
                     : "memory");
        tmem_wait_ld();
#if FW_OPT_ACC2
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const unsigned long long r2 = f2_add(f2_pack(acc[c + i], acc[c + i + 1]),
                                               f2_pack(__uint_as_float(u[i]), __uint_as_float(u[i + 1])));
          f2_unpack(r2, acc[c + i], acc[c + i + 1]);
        }
#else
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c + i] += __uint_as_float(u[i]);
#endif
      }
    }
#endif
    tc_fence_before_sync();
    mbar_arrive(&bars->d_free[db]);
    if (trace) FW_EVT(2, 16 + q, 2);
  }
}

// reciprocal square root to fp32 rounding: MUFU seed + one Newton step (the IEEE sqrt / div sequences of the generic
// feature code are ~100 dependent cycles each, and an entry's latency is what bounds the geometry role)
__device__ __forceinline__ float fw_rsqrt(float a) {
  const float y = rsqrtf(a);
  return y * fmaf(-0.5f * a, y * y, 1.5f);
}
// dihedral as [cos, sin] (reference ann.py:338-351): C = n1.n2, S = (n1.r34) |r23|, out = (C, S) / sqrt(C^2 + S^2)
template <class Out>
__device__ __forceinline__ void fw_dihedral_cos_sin(V3 x0, V3 x1, V3 x2, V3 x3, int off, Out& out) {
  const V3 r12 = sub(x1, x0), r23 = sub(x2, x1), r34 = sub(x3, x2);
  const V3 n1 = cross(r12, r23), n2 = cross(r23, r34);
  const float d23 = dot(r23, r23);
  const float C = dot(n1, n2);
  const float S = dot(n1, r34) * (d23 * fw_rsqrt(d23));
  const float ir = fw_rsqrt(fmaf(C, C, S * S));
  out(off, C * ir);
  out(off + 1, S * ir);
}

// A frame is copied from the 16-byte boundary below it, rounded up to 16 bytes: up to 15 bytes past its end.  That is
// inside x for every frame but the last one of the batch, which is staged only if it ends on the grid.
__device__ __forceinline__ bool fw_frame_is_staged(long long f, long long L, uint32_t off, uint32_t fbytes) {
  return f + 1 < L || ((off + fbytes) & 15u) == 0u;
}

// plan tables: shared memory when they fit (TS), else global memory
template <bool TS>
__device__ __forceinline__ int fw_tab_i(const unsigned char* smem, int off, const int* g, int i) {
  return TS ? reinterpret_cast<const int*>(smem + off)[i] : __ldg(g + i);
}
template <bool TS>
__device__ __forceinline__ float fw_tab_f(const unsigned char* smem, int off, const float* g, int i) {
  return TS ? reinterpret_cast<const float*>(smem + off)[i] : __ldg(g + i);
}

// ================= geometry role: moments, raw position atoms, invariant features -> scratch =================
// The FW_NGW warps of a group work on ONE frame at a time (thread = alignment atom / position atom / invariant entry),
// so the frames behind it in the ring are pure prefetch and a frame's latency is that of one entry, not of
// n_entries / 32 iterations of a single warp.  Alternate frames go to alternate groups.
// The warps never meet: each writes its results STRAIGHT into the frame's scratch row (small L2 stores, no staging
// row, no copy-out pass) and its own 12-float moment partial into the row header -- the converter adds the
// partials.  The per-frame group barrier, the header pass and the staging copy were 40 % of a frame's latency
// (tests/cuda/fw_trace.cu), and a frame's latency is what bounds this role.  Every frame is staged, so all
// coordinate loads are shared-memory loads with 32-bit addresses (the role lives on 56 registers).
template <bool TS>
__device__ __forceinline__ void fw_geometry_role(const FwParams& P, FwBars* bars, unsigned char* smem, long long L,
                                                 int ntile_cta, int gwarp, int lane) {
  const int grp = gwarp / FW_NGW;
  const int gw = gwarp % FW_NGW;
  const int gt = gw * 32 + lane;                    // thread within the group
  constexpr int gstride = FW_NGW * 32;
  const bool aligned = P.n_align > 0;
  const unsigned nfr = (unsigned)ntile_cta * FW_M;
  float* const cta_scratch = P.scratch + (long long)blockIdx.x * P.cta_floats;
  unsigned long long pol_keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  int slot = 0, rs = grp % P.n_ring, prev_slot = -1;
  uint32_t spar = 0, rpar = (uint32_t)(grp / P.n_ring) & 1u;
  unsigned k2 = 0;                                   // this group's frame counter
  FW_TRACE_ONLY(long long tw_slot = 0; long long tw_frame = 0; long long tw_bar = 0;)
#pragma unroll 1
  for (unsigned ic = (unsigned)grp; ic < nfr; ic += FW_NGG, ++k2) {
    const unsigned it = ic / FW_M;
    const int r = (int)(ic & (FW_M - 1));
    if (gw == 0 && grp == 0 && r == 0) {
      FW_TILE(0, it, 0, clock64());
      FW_TILE(0, it, 1, tw_slot);
      FW_TILE(0, it, 2, tw_frame);
      FW_TILE(0, it, 3, tw_bar);
    }
    const long long f = ((long long)blockIdx.x + (long long)it * gridDim.x) * FW_M + r;
    const int rr = r & (FW_SUB - 1);
    if (gw == 0 && grp == 0) FW_EVT(0, k2, 0);
    if (rr < FW_NGG) {                              // this group's first frame of a scratch sub-tile: slot must be free
      if (ic >= FW_NGG && ++slot == P.n_slots) {
        slot = 0;
        spar ^= 1u;
      }
      {
        // every lane waits (not lane 0 + __syncwarp): after a one-lane branch with a wait loop inside the warp stayed
        // SPLIT (lane 0 | lanes 1-31) for the rest of the frame -- every instruction issued twice, every shuffle through
        // the WARPSYNC.COLLECTIVE slow path (profiles/r3h: 13 collectives x 8 per frame)
        FW_WAIT_BEGIN();
        mbar_wait_hint(&bars->s_free[slot], spar ^ 1u);
        FW_WAIT_END(tw_slot);
      }
    }
    if (f < L) {
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 1);
      {
        FW_WAIT_BEGIN();
        mbar_wait_hint(&bars->x_full[rs], rpar);
        FW_WAIT_END(tw_frame);
      }
      // the frame sits `off` bytes into its ring slot (copies start on the 16-byte grid)
      const uint32_t off = (uint32_t)((unsigned long long)f * (unsigned long long)(12 * P.n_inp) +
                                      (unsigned long long)reinterpret_cast<uintptr_t>(P.x_base)) & 15u;
      const float* xf = reinterpret_cast<const float*>(smem + P.off_ring + (size_t)rs * P.ring_slot_bytes + off);
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 2);
      float* const row = cta_scratch + (long long)slot * P.slot_floats + (long long)rr * P.row_floats;
      float* const units = row + FW_HDR_FLOATS;
      // Order of a frame: every READ of the staged frame first (moment loop, position atoms, the entry's atoms into
      // registers), then the ring slot goes back to the X producer, then the dihedral arithmetic and its stores, which
      // need the frame no more.  A group gets a frame per max(T_frame, T_load), and the ring holds only four frames of
      // a C3 system: the earlier a slot is released, the more of the reload overlaps the group's own work.
      float pvx = 0.f, pvy = 0.f, pvz = 0.f;
      if (aligned) {
        // pivoted one-pass moments (reference ann.py:179-187): d_k = x_k - x_{A_0}, H = sum d_k^T y_k (the
        // reference is centred); this warp's partial of H and of sum d_k goes to the row header
        const float* p0 = xf + 3 * fw_tab_i<TS>(smem, P.off_aidx, P.align_idx, 0);
        pvx = p0[0]; pvy = p0[1]; pvz = p0[2];
        float m[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) m[i] = 0.f;
#pragma unroll 1
#ifdef FW_DBG_SKIP_MOM
        for (int k = gt; k < 32; k += gstride) {
#else
        for (int k = gt; k < P.n_align; k += gstride) {
#endif
          const float* p = xf + 3 * fw_tab_i<TS>(smem, P.off_aidx, P.align_idx, k);
          const float px = p[0] - pvx, py = p[1] - pvy, pz = p[2] - pvz;
          const float y0 = fw_tab_f<TS>(smem, P.off_ref, P.ref_x, 3 * k), y1 = fw_tab_f<TS>(smem, P.off_ref, P.ref_x, 3 * k + 1),
                      y2 = fw_tab_f<TS>(smem, P.off_ref, P.ref_x, 3 * k + 2);
          m[0] = fmaf(px, y0, m[0]); m[1] = fmaf(px, y1, m[1]); m[2] = fmaf(px, y2, m[2]);
          m[3] = fmaf(py, y0, m[3]); m[4] = fmaf(py, y1, m[4]); m[5] = fmaf(py, y2, m[5]);
          m[6] = fmaf(pz, y0, m[6]); m[7] = fmaf(pz, y1, m[7]); m[8] = fmaf(pz, y2, m[8]);
          m[9] += px; m[10] += py; m[11] += pz;
          if (P.pos_is_align) {                     // the same atom is position unit k: its raw coordinates are here
            float* dst = units + 4 * k;
            fw_st2(dst, px, py, pol_keep);
            fw_st1(dst + 2, pz, pol_keep);
            if (k >= P.n_inv) fw_st1(dst + 3, 0.f, pol_keep);
          }
        }
        // the loop's trip count differs between lanes: without this the two halves of the warp ran the rest of the frame
        // one after the other (every instruction issued twice, each shuffle wrapped in a WARPSYNC.COLLECTIVE call)
        __syncwarp();
        sb_reduce12_store(m, row + gw * 12, lane);
      }
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 3);
      // raw (pivot-relative) coordinates of the position atoms; the w slot of a unit belongs to invariant
      // column u when there is one, else it is zero
      if (!(aligned && P.pos_is_align)) {
        for (int u = gt; u < P.n_pos; u += gstride) {
          const float* p = xf + 3 * fw_tab_i<TS>(smem, P.off_pos, P.pos_atom, u);
          float* dst = units + 4 * u;
          fw_st2(dst, p[0] - pvx, p[1] - pvy, pol_keep);
          fw_st1(dst + 2, p[2] - pvz, pol_keep);
          if (u >= P.n_inv) fw_st1(dst + 3, 0.f, pol_keep);
        }
        __syncwarp();
      }
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 4);
      // invariant features (bond / angle / dihedral, ann.py:323-351) from the raw coordinates.  With at most one
      // entry per thread (C3: 100 dihedrals on 128 threads) a [cos, sin] dihedral only LOADS its four atoms here
      FwInvOut out{units, P.n_pos, pol_keep};
      bool late_dihedral = false;
      V3 d0, d1, d2, d3;
      int d_off = 0;
      {
        Rigid none;
        const int* ient = TS ? reinterpret_cast<const int*>(smem + P.off_ent) : P.inv_ent;
        if (P.n_inv_ent <= gstride) {
          if (gt < P.n_inv_ent) {
            const Entry en = load_entry(ient + ENTRY_INTS * gt);
            if (FW_OPT_LATE && en.type == FEAT_DIHEDRAL && !P.use_angle) {
              d0 = ld3(xf, en.a0); d1 = ld3(xf, en.a1); d2 = ld3(xf, en.a2); d3 = ld3(xf, en.a3);
              d_off = en.off;
              late_dihedral = true;
            } else {
              feature_forward(en, xf, false, none, P.use_angle, out);
            }
          }
        } else {
          for (int e = gt; e < P.n_inv_ent; e += gstride) {
            const Entry en = load_entry(ient + ENTRY_INTS * e);
            if (en.type == FEAT_DIHEDRAL && !P.use_angle)
              fw_dihedral_cos_sin(ld3(xf, en.a0), ld3(xf, en.a1), ld3(xf, en.a2), ld3(xf, en.a3), en.off, out);
            else feature_forward(en, xf, false, none, P.use_angle, out);
          }
        }
      }
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 5);
      __syncwarp();                                 // every lane has read what it needs of the frame
      if (lane == 0) {
        mbar_arrive(&bars->x_empty[rs]);
        // publish this warp's part of the PREVIOUS frame's row: those stores were issued a whole frame ago, so
        // the release does not stall
        if (prev_slot >= 0) mbar_arrive(&bars->s_full[prev_slot]);
      }
#ifdef FW_DBG_SKIP_DIH
      if (late_dihedral) { out(d_off, d0.x); out(d_off + 1, d3.y); }
#else
      if (late_dihedral) fw_dihedral_cos_sin(d0, d1, d2, d3, d_off, out);
#endif
      if (gt == 0 && P.n_inv > P.n_pos) {           // zero the unused tail of the last invariant unit
        for (int v = P.n_inv; ((v - P.n_pos) & 3) != 0; ++v) out(v, 0.f);
      }
      prev_slot = slot;
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 6);
      rs += FW_NGG;
      while (rs >= P.n_ring) {
        rs -= P.n_ring;
        rpar ^= 1u;
      }
      if (gw == 0 && grp == 0) FW_EVT(0, k2, 7);
    } else if (lane == 0) {
      mbar_arrive(&bars->s_full[slot]);             // past the end of the batch: keep the sub-tile's count whole
    }
  }
  __syncwarp();
  if (lane == 0 && prev_slot >= 0) mbar_arrive(&bars->s_full[prev_slot]);
}

// STORE_H: also leave the hidden activations behind (value-and-gradient path).  A template parameter, not a run-time
// test: the unused store code alone cost the plain forward 5 % (the kernel is sensitive to its instruction footprint).
template <int ACT, int KU, bool STORE_H = false>
__global__ void __launch_bounds__(FW_THREADS, 1)
fused_wide_forward_kernel(const __grid_constant__ FwParams P, const float* __restrict__ x, float* __restrict__ y,
                          long long L) {
  constexpr int FW_KU = KU, FW_KC = fw_kc(KU), FW_SEGC = FW_SEG_K / FW_KC;      // chunks per accumulation segment
  constexpr int FW_A_HALF = fw_a_half(KU), FW_STAGE_BYTES = fw_stage_bytes(KU), FW_CONV_CHUNK = fw_conv_chunk(KU);
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
      mbar_init(&bars->x_full[s], FW_OPT_XPROD2 ? 64 : 32);   // every lane of the X producer warps arrives (cp.async completion)
      mbar_init(&bars->x_empty[s], FW_NGW);   // every warp of the frame's group releases it
    }
    for (int s = 0; s < FW_MAX_SLOTS; ++s) {
      mbar_init(&bars->s_full[s], FW_SUB * (FW_NGW));   // each warp publishes its part of a row
      mbar_init(&bars->s_free[s], 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&bars->turn[s], 128);
    mbar_init(&bars->l2_full, 1);
    mbar_init(&bars->l2_free, 16 * 32);
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

  // Nothing long-lived is computed here: every role derives what it needs from the (constant-bank) parameters itself.
  // Values that stay live across the role dispatch are spilled under the small per-role register budgets, and with
  // 200+ KB of shared memory configured there is next to no L1 left: a spill reload costs an L2 round trip
  // (tests/cuda/fw_trace.cu: 15 k cycles for the 7-iteration moment loop of a frame instead of 1.3 k).
#define FW_NTILES ((int)((L + FW_M - 1) / FW_M))
#define FW_NCHUNKS_TILE (P.nkc1 + (P.n_hidden == 2 ? P.nkc2 : 0))
#define FW_STAGES (smem + P.off_stage)
#define FW_RING (smem + P.off_ring)
#define FW_CTA_SCRATCH (P.scratch + (long long)blockIdx.x * P.cta_floats)

  // every role states its register budget at the TOP of its own branch: ptxas compiles a region under the smallest
  // setmaxnreg that can reach it, and it does not correlate two tests of `warp` (a budget set in a separate if-chain
  // ahead of the dispatch put the geometry role on the control warps' 24 registers: 57 spill stores per frame loop)
  if (warp >= FW_W_WPROD && warp < FW_W_GEO) {
  // ---- the control warpgroup (W producer, MMA issuer, two X producers): ONE setmaxnreg for its four warps (the
  // instruction is .aligned per warpgroup: four copies in four branches are an illegal instruction)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_CTRL));
  if (warp == FW_W_XPROD || (FW_OPT_XPROD2 && warp == FW_W_XPROD + 1)) {
    // ================= X producers: whole frames into the shared-memory ring =================
    // 16-byte cp.async (LDGSTS) copies, NOT the bulk-copy engine: that engine is one in-order queue per SM, and with
    // 24 KB frame loads (HBM latency) always in flight every small bulk store and every fence.proxy.async of the
    // other roles waited 4.6 k cycles behind them (tests/cuda/fw_trace.cu).  A copy starts at the 16-byte boundary
    // below the frame and may run up to 15 bytes past it -- never past the end of x (fw_frame_is_staged: the one
    // frame of a batch that would is copied float by float).
    // Two warps take alternate 2 KB pieces of every frame, four copies per loop trip: with one warp and a rolled loop
    // the ISSUE of a frame's 47 copies took ~850 cycles, a quarter of the time from "slot free" to "frame landed",
    // and that time, not the geometry, bounded a group's frame rate (ring of four frames, two groups).
    const int half = warp - FW_W_XPROD;
    constexpr uint32_t nprod = FW_OPT_XPROD2 ? 2 : 1;
    int rs = 0;
    uint32_t rpar = 0;
    const uint32_t fbytes = 12u * (uint32_t)P.n_inp;
    const int ntiles = FW_NTILES;
    // x is read exactly once: mark its lines evict-first so that the stream does not push the CTA's transposition
    // scratch (re-used every tile) out of L2
    unsigned long long pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long f0 = (long long)tile * FW_M;
      const int nf = L - f0 < FW_M ? (int)(L - f0) : FW_M;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(x) + f0 * (long long)fbytes;
#pragma unroll 1
      for (int r = 0; r < nf; ++r, src += fbytes) {
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
#ifdef FW_DBG_HALF_X
        const uint32_t bytes = ((fbytes / 2) + off + 15u) & ~15u;
#else
        const uint32_t bytes = (fbytes + off + 15u) & ~15u;
#endif
        const uint32_t d0 = smem_u32(FW_RING) + (uint32_t)rs * (uint32_t)P.ring_slot_bytes;
        mbar_wait_hint(&bars->x_empty[rs], rpar ^ 1u);
        const unsigned char* s0 = src - off;
        uint32_t o = (uint32_t)half * FW_PIECE + (uint32_t)lane * 16u;
        // the batch's very last frame goes float by float when its 16-byte copies would run past the end of x
        if (!fw_frame_is_staged(f0 + r, L, off, fbytes)) {
          for (uint32_t b = (uint32_t)(half * 32 + lane) * 4u; b < fbytes; b += nprod * 32u * 4u)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + off + b), "l"(src + b) : "memory");
          o = bytes;
        }
        for (; o + 3u * 512u < bytes; o += nprod * FW_PIECE) {                                   // a whole 2 KB piece
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d0 + o + q * 512u),
                         "l"(s0 + o + q * 512u), "l"(pol_stream)
                         : "memory");
        }
        for (int q = 0; q < 4; ++q)                                                               // the ragged last piece
          if (o + q * 512u < bytes)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d0 + o + q * 512u),
                         "l"(s0 + o + q * 512u), "l"(pol_stream)
                         : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars->x_full[rs]))
                     : "memory");
        if (++rs == P.n_ring) {
          rs = 0;
          rpar ^= 1u;
        }
      }
    }
  } else if (warp == FW_W_WPROD) {
    // ================= W producer: pre-packed weight blocks, layer 1 then layer 2 =================
    if (lane == 0) {
      const uint32_t bytes1 = 2u * FW_KC * (uint32_t)P.n1p * 4u, bytes2 = 2u * FW_KC * (uint32_t)P.n2p * 4u;
      int s = 0;
      uint32_t par = 0;
      const int ntiles = FW_NTILES;
      const int nchunks_tile = FW_NCHUNKS_TILE;
      unsigned char* const stages = FW_STAGES;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
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
    unsigned sg = 0, it = 0;
    const int stride2 = P.n2p <= 128 ? 128 : 256;            // layer 2: 4 (or 2) accumulators, no drain in between
    const int cpa2 = (P.nkc2 + (512 / stride2) - 1) / (512 / stride2);
    const int ntiles = FW_NTILES;
    unsigned char* const stages = FW_STAGES;
    FW_TRACE_ONLY(long long tw_a = 0; long long tw_b = 0; long long tw_d = 0;)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      FW_TILE(3, it, 0, clock64());
      FW_TILE(3, it, 1, tw_a);
      FW_TILE(3, it, 2, tw_b);
      FW_TILE(3, it, 3, tw_d);
      if (it > 0 && P.n_hidden == 2) {                       // layer 2's accumulators of the previous tile are read
        mbar_wait_hint(&bars->l2_free, (it - 1u) & 1u);
        tc_fence_after_sync();
      }
      for (int layer = 0; layer < P.n_hidden; ++layer) {
        const int nkc = layer == 0 ? P.nkc1 : P.nkc2;
        const int np = layer == 0 ? P.n1p : P.n2p;
        const uint32_t idesc = idesc_tf32(FW_M, np);
        const uint32_t lbo_b = (uint32_t)np * 16u;
        if (layer == 1) {                                    // every layer-1 segment has been summed by the epilogue
          if (sg >= 2) mbar_wait_hint(&bars->d_free[(sg - 2u) & 1u], ((sg - 2u) >> 1) & 1u);
          mbar_wait_hint(&bars->d_free[(sg - 1u) & 1u], ((sg - 1u) >> 1) & 1u);
          tc_fence_after_sync();
        }
        for (int kc = 0; kc < nkc; ++kc) {
          uint32_t d;
          bool first, seg_last = false;
          const int tci = (int)it * 128 + layer * 64 + kc;     // trace index
          (void)tci;
          FW_EVT(3, tci, 0);
          if (layer == 0) {
            const int db = (int)(sg & 1u);
            d = (uint32_t)db * FW_NMAX;
            first = (kc % FW_SEGC) == 0;
            seg_last = (kc % FW_SEGC) == FW_SEGC - 1 || kc == nkc - 1;
            if (first) {
              FW_WAIT_BEGIN();
              mbar_wait_hint(&bars->d_free[db], ((sg >> 1) & 1u) ^ 1u);
              FW_WAIT_END(tw_d);
              tc_fence_after_sync();
            }
          } else {
            d = (uint32_t)((kc / cpa2) * stride2);
            first = (kc % cpa2) == 0;
          }
          FW_EVT(3, tci, 1);
          {
            FW_WAIT_BEGIN();
            mbar_wait_hint(&bars->a_full[s], par);
            FW_WAIT_END(tw_a);
          }
          FW_EVT(3, tci, 2);
          {
            FW_WAIT_BEGIN();
            mbar_wait_hint(&bars->b_full[s], par);
            FW_WAIT_END(tw_b);
          }
          FW_EVT(3, tci, 3);
          if (layer == 1 && kc == 0) FW_TILE(3, it, 4, clock64());
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
#ifdef FW_DBG_SKIP_CROSS      // timing probe only (wrong results): how much of the period is MMA count / operand fetch?
              mma_tf32_ss(d, al, bh, idesc, (!first || jj > 0) ? 1u : 0u);
#else
              mma_tf32_ss(d, al, bh, idesc, (!first || jj > 0) ? 1u : 0u);
              mma_tf32_ss(d, ah, bl, idesc, 1);
#endif
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
            if (leader) mma_commit(&bars->d_full[sg & 1u]);
            ++sg;
          }
          if (layer == 1 && kc == nkc - 1 && leader) mma_commit(&bars->l2_full);
          __syncwarp();
          FW_EVT(3, tci, 4);
          fw_stage_step(s, par, P.n_stages);
        }
      }
    }
  }
  } else if (warp >= FW_W_GEO) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_GEO));
    const int ntile_cta = (FW_NTILES - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool tabs = P.off_ent >= 0 && (P.n_align == 0 || P.off_aidx >= 0) && (P.off_pos >= 0 || P.pos_is_align);
    if (tabs) fw_geometry_role<true>(P, bars, smem, L, ntile_cta, warp - FW_W_GEO, lane);
    else fw_geometry_role<false>(P, bars, smem, L, ntile_cta, warp - FW_W_GEO, lane);
  } else if (warp < FW_W_EPI) {
    // ================= converter: rotation per frame, then K-chunks scratch -> TF32 hi / lo operand tiles ==========
    // The scratch is read with cp.async into a small staging ring, P.conv_depth chunks ahead: a read of the L2 scratch
    // takes 3-4 k cycles while the frame stream keeps the memory pipe busy (tests/cuda/fw_trace.cu), and with one
    // chunk of register prefetch that latency was the period of the whole MMA chain.
    if (FW_REGS_CONV > FW_REGS_LAUNCH) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FW_REGS_CONV));
    if (FW_REGS_CONV < FW_REGS_LAUNCH) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_CONV));
    const int row = tid;                               // 0 .. 127
    const bool aligned = P.n_align > 0;
    constexpr int ngw = FW_NGW;
    const int depth = P.conv_depth;
    unsigned char* const cst = smem + P.off_cstage + row * 16;     // this thread's 16-byte column of the staging ring
    unsigned it = 0;
    int s = 0;
    uint32_t par = 0;
    const int ntiles = FW_NTILES;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      if (warp == 0) FW_EVT(1, it, 0);
      if (warp == 0) FW_TILE(1, it, 0, clock64());
      const unsigned j = it * (FW_M / FW_SUB) + (unsigned)warp;
      const unsigned use = j / (unsigned)P.n_slots;
      const int slot = (int)(j - use * (unsigned)P.n_slots);
      mbar_wait_hint(&bars->s_full[slot], use & 1u);
      if (warp == 0) FW_EVT(1, it, 1);
      if (warp == 0) FW_TILE(1, it, 1, clock64());
      const float* myrow = FW_CTA_SCRATCH + (long long)slot * P.slot_floats + (long long)lane * P.row_floats;
      const float* units = myrow + FW_HDR_FLOATS;
      const bool valid = (long long)tile * FW_M + row < L;
      // raw chunk kc of this row -> staging slot kc mod depth (one commit group per chunk, empty ones included, so
      // that "all but the newest depth - 1 groups" always means "chunk kc has landed")
      auto fetch = [&](int kc) {
        if (kc < P.nkc1 && valid) {
          const uint32_t d0 = smem_u32(cst) + (uint32_t)(kc % depth) * FW_CONV_CHUNK;
          const float* src = units + (size_t)kc * FW_KC;
#pragma unroll
          for (int q = 0; q < FW_KU; ++q)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)q * (FW_M * 16)), "l"(src + 4 * q)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      for (int kc = 0; kc < depth - 1; ++kc) fetch(kc);
      Rigid rg;
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (aligned) {
        // moment partials of the frame's geometry warps, added in a fixed order: ((p0 + p1) + (p2 + p3)) [+ ...]
        float hs[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) hs[i] = 0.f;
        if (valid) {
          const float4* h = reinterpret_cast<const float4*>(myrow);
          for (int q4 = 0; q4 < ngw; q4 += 4) {
            float quad[12];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              float a[12];
#pragma unroll
              for (int v4 = 0; v4 < 6; ++v4) {           // two partials = 24 floats = 6 float4
                const float4 t = __ldcg(h + (q4 + 2 * pr) * 3 + v4);
                const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int i = 4 * v4 + c;
                  if (i < 12) a[i] = tv[c];
                  else a[i - 12] += tv[c];
                }
              }
#pragma unroll
              for (int i = 0; i < 12; ++i) quad[i] = pr == 0 ? a[i] : quad[i] + a[i];
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) hs[i] = q4 == 0 ? quad[i] : hs[i] + quad[i];
          }
        } else {
          hs[0] = hs[4] = hs[8] = 1.f;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) rg.H[i] = hs[i];
        const float inv_n = 1.0f / (float)P.n_align;
        c0 = hs[9] * inv_n; c1 = hs[10] * inv_n; c2 = hs[11] * inv_n;
        kabsch_rotation(rg);                          // reference ann.py:188-195 as a quaternion eigenproblem
      }
      if (warp == 0) FW_EVT(1, it, 2);
      if (warp == 0) FW_TILE(1, it, 2, clock64());
      // layer 2's operand chunks of the previous tile share the stage ring: they must all have been written before
      // this role asks for a stage again (a waiter may be at most one phase ahead of an mbarrier)
      if (it > 0 && P.n_hidden == 2) mbar_wait_hint(&bars->turn[3], (it - 1u) & 1u);
      if (warp == 0) FW_EVT(1, it, 3);
      if (warp == 0) FW_TILE(1, it, 3, clock64());
      FW_TRACE_ONLY(long long tw_empty = 0;)
      for (int kc = 0; kc < P.nkc1; ++kc) {
        fetch(kc + depth - 1);
        if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else asm volatile("cp.async.wait_group 3;" ::: "memory");
        float4 v[FW_KU];
        {
          const unsigned char* src = cst + (size_t)(kc % depth) * FW_CONV_CHUNK;
#pragma unroll
          for (int q = 0; q < FW_KU; ++q)
            v[q] = valid ? *reinterpret_cast<const float4*>(src + q * (FW_M * 16)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < FW_KU; ++q) {
          const int u = kc * FW_KU + q;
          if (u >= P.n_units) {
            v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef FW_DBG_SKIP_ROT
          } else if (aligned && u < P.n_pos) {        // z = (p - c) R  (ann.py:197), p and c relative to the pivot
            const float dx = v[q].x - c0, dy = v[q].y - c1, dz = v[q].z - c2;
            v[q].x = fmaf(dx, rg.R[0], fmaf(dy, rg.R[3], dz * rg.R[6]));
            v[q].y = fmaf(dx, rg.R[1], fmaf(dy, rg.R[4], dz * rg.R[7]));
            v[q].z = fmaf(dx, rg.R[2], fmaf(dy, rg.R[5], dz * rg.R[8]));
#endif
          }
        }
        if (warp == 0 && it == 0) FW_EVT(1, 16 + kc, 0);
        {
          FW_WAIT_BEGIN();
          mbar_wait_hint(&bars->empty[s], par ^ 1u);
          FW_WAIT_END(tw_empty);
        }
        if (warp == 0 && it == 0) FW_EVT(1, 16 + kc, 1);
        fw_store_units<KU>(FW_STAGES + (size_t)s * FW_STAGE_BYTES, row, v);
        if (warp == 0 && it == 0) FW_EVT(1, 16 + kc, 2);
        fence_proxy_async_smem();
        mbar_arrive(&bars->a_full[s]);
        if (warp == 0 && it == 0) FW_EVT(1, 16 + kc, 3);
        fw_stage_step(s, par, P.n_stages);
      }
      // the row is dead: drop its L2 lines instead of letting them be written back to DRAM
#if FW_OPT_POLICY
      asm volatile("cp.async.wait_all;" ::: "memory");
      for (int b = 0; b < P.row_floats; b += 32) asm volatile("discard.global.L2 [%0], 128;" ::"l"(myrow + b) : "memory");
#endif
      __syncwarp();
      if (warp == 0) FW_EVT(1, it, 4);
      if (warp == 0) FW_TILE(1, it, 4, clock64());
      if (warp == 0) FW_TILE(1, it, 5, tw_empty);
      if (lane == 0) mbar_arrive(&bars->s_free[slot]);
      if (P.n_hidden == 2) fw_stage_skip(s, par, P.nkc2, P.n_stages);
    }
  } else if (warp < FW_W_WPROD) {
    // ================= epilogue: segment sums, activations, layer 2 operand chunks, last layer, y =================
    if (FW_REGS_EPI > FW_REGS_LAUNCH) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FW_REGS_EPI));
    const int e = (warp - FW_W_EPI) >> 2;            // this warpgroup owns accumulator columns [64 e, 64 e + 64)
    const int row = tid & 127;
    const int col0 = FW_CW * e;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    const float* b1 = reinterpret_cast<const float*>(smem + P.off_b1);
    const float* b2 = reinterpret_cast<const float*>(smem + P.off_b2);
    const float* w3 = reinterpret_cast<const float*>(smem + P.off_w3);
    float* ypart = reinterpret_cast<float*>(smem + P.off_ypart);
    unsigned sg = 0, it = 0;
    const int ntiles = FW_NTILES;
    const int nseg1 = (P.nkc1 + FW_SEGC - 1) / FW_SEGC;
    unsigned char* const stages = FW_STAGES;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      float acc[FW_CW];
      if (warp == FW_W_EPI) FW_EVT(2, it, 0);
      if (warp == FW_W_EPI) FW_TILE(2, it, 0, clock64());
      fw_sum_segments(acc, nseg1, col0, P.n1p, lane_base, bars, sg, warp == FW_W_EPI && it == 0);
      if (warp == FW_W_EPI) FW_EVT(2, it, 1);
      if (warp == FW_W_EPI) FW_TILE(2, it, 1, clock64());
#pragma unroll
      for (int c = 0; c < FW_CW; ++c)
        if (col0 + c < P.n1p) acc[c] = fw_act<ACT>(acc[c] + b1[col0 + c]);
      if (STORE_H && P.h1_out != nullptr && (long long)tile * FW_M + row < L) {       // value-and-gradient: keep h1 for act'
        float* dst = P.h1_out + ((long long)tile * FW_M + row) * P.n1 + col0;
        if ((P.n1 & 3) == 0) {
#pragma unroll
          for (int c = 0; c < FW_CW; c += 4)
            if (col0 + c < P.n1) __stcg(reinterpret_cast<float4*>(dst + c), make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
        } else {
#pragma unroll
          for (int c = 0; c < FW_CW; ++c)
            if (col0 + c < P.n1) __stcg(dst + c, acc[c]);
        }
      }
      int nlast = P.n1p, lcol0 = col0, lcw = FW_CW;      // columns of the last hidden layer this thread holds
      if (P.n_hidden == 2) {
        // h1 goes back into the operand ring as layer 2's A chunks, warpgroup after warpgroup (chunk order)
        if (e > 0) mbar_wait_hint(&bars->turn[e - 1], it & 1u);
        const unsigned gbase = it * (unsigned)FW_NCHUNKS_TILE + (unsigned)P.nkc1;
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
            fw_store_units<KU>(stages + (size_t)s * FW_STAGE_BYTES, row, v);
            fence_proxy_async_smem();
            mbar_arrive(&bars->a_full[s]);
          }
        }
        mbar_arrive(&bars->turn[e]);
        if (warp == FW_W_EPI) FW_EVT(2, it, 2);
        // layer 2: sum of its accumulators (fixed order), this thread's stride2 / 4 columns
        const int stride2 = P.n2p <= 128 ? 128 : 256;
        const int cw2 = stride2 / 4;
        lcol0 = cw2 * e;
        lcw = cw2;
#pragma unroll
        for (int i = 0; i < FW_CW; ++i) acc[i] = 0.f;
        mbar_wait_hint(&bars->l2_full, it & 1u);
        tc_fence_after_sync();
        if (warp == FW_W_EPI) FW_EVT(2, it, 3);
        if (warp == FW_W_EPI) FW_TILE(2, it, 3, clock64());
        const int cpa2 = (P.nkc2 + (512 / stride2) - 1) / (512 / stride2);
        const int nacc2 = (P.nkc2 + cpa2 - 1) / cpa2;
        for (int a = 0; a < nacc2; ++a) {
#pragma unroll
          for (int c = 0; c < FW_CW; c += 8) {
            if (c < cw2 && lcol0 + c < P.n2p) {
              uint32_t u[8];
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                           : "r"(lane_base + (uint32_t)(a * stride2 + lcol0 + c))
                           : "memory");
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[c + i] += __uint_as_float(u[i]);
            }
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&bars->l2_free);
#pragma unroll
        for (int c = 0; c < FW_CW; ++c)
          if (c < cw2 && lcol0 + c < P.n2p) acc[c] = fw_act<ACT>(acc[c] + b2[lcol0 + c]);
        if (STORE_H && P.h2_out != nullptr && (long long)tile * FW_M + row < L) {
          float* dst = P.h2_out + ((long long)tile * FW_M + row) * P.n2 + lcol0;
          if ((P.n2 & 3) == 0) {
#pragma unroll
            for (int c = 0; c < FW_CW; c += 4)
              if (c < cw2 && lcol0 + c < P.n2)
                __stcg(reinterpret_cast<float4*>(dst + c), make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
          } else {
#pragma unroll
            for (int c = 0; c < FW_CW; ++c)
              if (c < cw2 && lcol0 + c < P.n2) __stcg(dst + c, acc[c]);
          }
        }
        nlast = P.n2p;
      }
      // last (narrow) layer: partial dot products over this thread's columns, summed in a fixed order
      for (int o = 0; o < P.kout; ++o) {
        float pd = 0.f;
#pragma unroll
        for (int c = 0; c < FW_CW; ++c)
          if (c < lcw && lcol0 + c < nlast) pd = fmaf(acc[c], w3[o * P.nlastp + lcol0 + c], pd);
        ypart[(e * FW_M + row) * P.kout + o] = pd;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(16 * 32) : "memory");
      if (e == 0) {
        const long long f = (long long)tile * FW_M + row;
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
      if (warp == FW_W_EPI) FW_EVT(2, it, 4);
      if (warp == FW_W_EPI) FW_TILE(2, it, 4, clock64());
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512u);
#undef FW_NTILES
#undef FW_NCHUNKS_TILE
#undef FW_STAGES
#undef FW_RING
#undef FW_CTA_SCRATCH
}

// ---------------------------------------------------------------------------------------------------------------
// One-time packing (molann_b200_prepare): out[kc][hi | lo][(k / 4)][np][4] of  scale * W[n][colmap[k]]
// ---------------------------------------------------------------------------------------------------------------
__global__ void fw_pack_kernel(const float* __restrict__ W, int ldw, int N, const int* __restrict__ colmap, int Kp,
                               int np, float scale, float* __restrict__ out, int FW_KC) {
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
