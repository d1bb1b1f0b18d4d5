// staged_block.cuh -- preprocessing kernels for big systems (C3 / C5: thousands of atoms per frame) in which a
// whole CTA works on ONE frame at a time and the frames flow through a shared-memory ring.
//
// Why (tests/cuda/sb_trace.cu, profiles/r1_h): the warp-per-frame staged kernels (general.cuh) keep 4 (backward) or
// 8 (forward) frames per SM in flight and every warp walks its frame alone -- load, moments, rotation, 300 entries
// at 32 per step, scatter, store, strictly one after the other: 751 us for 32768 C3 frames against 258 us of HBM time.
// Here
//   * one elected thread keeps `stages` frames of bulk copies (x row [+ cotangent row]) in flight ahead of the math,
//   * the plan's constants (selection, reference, feature program) sit in shared memory,
//   * the work of a frame is split between two ROLES that run on different frames at the same time:
//       E (warps 1..7, one entry / alignment atom per thread): moments of frame j+1, then the entries of frame j
//         (rigid-motion invariant ones first, they do not need the rotation), backward: + 3x3 solve + scatter
//       G (warp 0): sum of the moment partials + rotation of frame j+1 (a 1700-cycle dependent chain)
//     handing (c, R, H) over through a double-buffered broadcast slot and two mbarrier pairs.  With every thread
//     waiting at CTA barriers for warp 0's rotation (first version) "barrier" was 39 % of all stalls
//     (profiles/r1_h) and the frame period was the SUM of the phases; now it is the longer of the two roles,
//   * the backward zero-fills the dense gradient row in HBM/L2 with plain 16-byte stores at the START of the frame and
//     adds the ~500 referenced atoms with fire-and-forget RED.ADD.F32 afterwards (both land in L2; DRAM sees one
//     write of the row).  A first version built the row in shared memory: fp32 shared-memory atomics are
//     ATOMS.CAST.SPIN loops (a load, an add and a compare-and-swap round trip each, 12 per dihedral) and were 40 %
//     of the frame's critical path, and the row buffer halved the CTAs per SM.
// Reductions are shuffle trees inside a warp and a fixed-order sum over the warps, so a frame's rigid transform does
// not depend on timing; the REDs of one phase commute as long as an atom collects at most two of them (C3 / C5),
// beyond that the row is reproducible to rounding.
//
// Tried and dropped (gpurun_out/c3_diag4.log): gathering only the referenced atoms with 4-byte cp.async into a compact
// buffer.  Shared memory per frame fell 4x, but every 12-byte atom costs a 64-byte DRAM access and nearly every
// 64-byte block of a C3 frame holds a referenced atom: forward 0.27 ms against 0.20 ms for the contiguous bulk copy.
#pragma once
#include "common.cuh"
#include "fused_ws.cuh"
#include "general.cuh"
#include "geometry.cuh"

namespace molann {

constexpr int SB_WARPS = 8;
constexpr int SB_THREADS = SB_WARPS * 32;
constexpr int SB_HEAD = 1536;                      // barriers + reduction scratch + broadcast slots

constexpr int SB_E_THREADS = SB_THREADS - 32;     // role E: warps 1..7
constexpr int SB_E_WARPS = SB_WARPS - 1;

struct SbSmem {
  unsigned long long x_full[4];                    // "row landed" per stage (transaction barriers)
  unsigned long long mom_full[2];                  // E -> G: moment partials of a frame are in red[k]
  unsigned long long rig_full[2];                  // G -> E: (c, R, H) of a frame are in bc[k]
  unsigned long long rig_empty[2];                 // E -> G: bc[k] has been consumed
  float red[2][SB_E_WARPS * 12 + 4];               // per-warp moment partials + pivot atom, double-buffered
  float bc[2][24];                                 // c[3] R[9] H[9], double-buffered
  float redm[SB_E_WARPS * 12];                     // backward: per-warp partials of M, sg
  float bcm[12];                                   // backward: dH[9] sg[3]
  int n_lead;                                      // leading position entries of an aligned plan
};
static_assert(sizeof(SbSmem) <= SB_HEAD, "header too small");

constexpr int SB_ZERO_BYTES = 4096;                // block of zeros the backward's row fill is bulk-stored from

struct SbLayout {                                  // byte offsets into dynamic shared memory (host-computed)
  int aidx_off, ref_off, ent_off;                  // plan constants
  int zero_off;                                    // backward: SB_ZERO_BYTES of zeros (-1: fill with plain stores)
  int ring_off;                                    // stages x [x row | cotangent row (backward)]
  int buf_bytes, fbuf_bytes, stages, total;
};

#ifdef MOLANN_WS_TRACE
__device__ long long g_sb_trace[64 * 16];
#define SB_EVT(it, ev)                                                                       \
  do {                                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 32 && (it) < 64) g_sb_trace[(it) * 16 + (ev)] = clock64(); \
  } while (0)
#else
#define SB_EVT(it, ev) \
  do {                 \
  } while (0)
#endif

// rows are copied from the 16-byte boundary below them, rounded up to 16 bytes
__device__ __forceinline__ bool sb_row_is_bulk(const float* base, long long f, long long L, int n) {
  const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(base + f * n) & 15u);
  return off == 0u || f + 1 < L;                   // the copy may run up to 15 bytes past the row: not on the last one
}
__device__ __forceinline__ uint32_t sb_row_bytes(const float* base, long long f, int n) {
  const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(base + f * n) & 15u);
  return ((uint32_t)n * 4u + off + 15u) & ~15u;
}
__device__ __forceinline__ const float* sb_row_ptr(const float* base, long long f, int n, unsigned char* slot) {
  const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(base + f * n) & 15u);
  return reinterpret_cast<const float*>(slot + off);
}

__device__ __forceinline__ void sb_stage_consts(const DevPlan& p, const SbLayout& lay, unsigned char* smem, SbSmem& s,
                                                int tid) {
  int* aidx = reinterpret_cast<int*>(smem + lay.aidx_off);
  float* ref = reinterpret_cast<float*>(smem + lay.ref_off);
  int* ent = reinterpret_cast<int*>(smem + lay.ent_off);
  for (int i = tid; i < p.n_align; i += SB_THREADS) aidx[i] = __ldg(p.align_idx + i);
  for (int i = tid; i < 3 * p.n_align; i += SB_THREADS) ref[i] = __ldg(p.ref_x + i);
  for (int i = tid; i < ENTRY_INTS * p.n_entries; i += SB_THREADS) ent[i] = __ldg(p.entries + i);
  if (tid == 0) {
    int nl = 0;
    if (p.n_align > 0)
      while (nl < p.n_entries && __ldg(p.entries + ENTRY_INTS * nl) == FEAT_POSITION) ++nl;
    s.n_lead = nl;
    for (int i = 0; i < lay.stages; ++i) mbar_init(&s.x_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.mom_full[i], 1);
      mbar_init(&s.rig_full[i], 1);
      mbar_init(&s.rig_empty[i], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
}

__device__ __forceinline__ void sb_e_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SB_E_THREADS) : "memory"); }

// Sum twelve per-lane values over the warp and store the twelve totals to dst[0..11].  A transposing butterfly:
// each step exchanges HALF of the values a lane still holds (12 -> 6 -> 3 -> 2 -> 1), 13 shuffles instead of the
// 60 of twelve separate butterflies -- the shuffle trees were 31 % of the instructions the first block kernel
// executed (profiles/r1_h) and SHFL shares the one-per-clock MIO port with shared-memory traffic.
__device__ __forceinline__ void sb_reduce12_store(const float (&v)[12], float* dst, int lane) {
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
  float w[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float keep = b16 ? v[6 + i] : v[i], send = b16 ? v[i] : v[6 + i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float u[4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float keep = b8 ? w[3 + i] : w[i], send = b8 ? w[i] : w[3 + i];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  u[3] = 0.f;
  float t[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b4 ? u[2 + i] : u[i], send = b4 ? u[i] : u[2 + i];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = b2 ? t[1] : t[0], send = b2 ? t[0] : t[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  const int sub = (b4 ? 2 : 0) + (b2 ? 1 : 0);
  if ((lane & 1) == 0 && sub < 3) dst[(b16 ? 6 : 0) + (b8 ? 3 : 0) + sub] = r;
}

// role E: moments of the alignment selection, one atom per thread -> per-warp partial sums in red[] (+ the pivot)
__device__ __forceinline__ void sb_moments_partial(const float* __restrict__ xf, const int* aidx, const float* ref,
                                                   int n_align, int et, int lane, int ewarp, float* red) {
  const float* p0 = xf + 3 * aidx[0];
  const float pv0 = p0[0], pv1 = p0[1], pv2 = p0[2];
  float v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = 0.f;
  for (int k = et; k < n_align; k += SB_E_THREADS) {
    const float* q = xf + 3 * aidx[k];
    const float px = q[0] - pv0, py = q[1] - pv1, pz = q[2] - pv2;
    const float y0 = ref[3 * k], y1 = ref[3 * k + 1], y2 = ref[3 * k + 2];
    v[0] = fmaf(px, y0, v[0]); v[1] = fmaf(px, y1, v[1]); v[2] = fmaf(px, y2, v[2]);
    v[3] = fmaf(py, y0, v[3]); v[4] = fmaf(py, y1, v[4]); v[5] = fmaf(py, y2, v[5]);
    v[6] = fmaf(pz, y0, v[6]); v[7] = fmaf(pz, y1, v[7]); v[8] = fmaf(pz, y2, v[8]);
    v[9] += px; v[10] += py; v[11] += pz;
  }
  sb_reduce12_store(v, red + ewarp * 12, lane);
  if (lane == 0 && ewarp == 0) {
    red[SB_E_WARPS * 12] = pv0; red[SB_E_WARPS * 12 + 1] = pv1; red[SB_E_WARPS * 12 + 2] = pv2;
  }
}
// role G: fixed-order sum over the E warps, rotation -> rg (identical in all 32 lanes)
__device__ __forceinline__ void sb_finish_rigid(int n_align, const float* red, Rigid& rg) {
  float v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < SB_E_WARPS; ++w) a += red[w * 12 + i];
    v[i] = a;
  }
  const float inv_n = 1.0f / (float)n_align;
#pragma unroll
  for (int i = 0; i < 9; ++i) rg.H[i] = v[i];
  rg.c[0] = fmaf(v[9], inv_n, red[SB_E_WARPS * 12]);
  rg.c[1] = fmaf(v[10], inv_n, red[SB_E_WARPS * 12 + 1]);
  rg.c[2] = fmaf(v[11], inv_n, red[SB_E_WARPS * 12 + 2]);
  kabsch_rotation(rg);
}
__device__ __forceinline__ void sb_store_rigid(const Rigid& rg, float* bc, int lane) {
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) bc[i] = rg.c[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) bc[3 + i] = rg.R[i];
  }
  if (lane == 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) bc[12 + i] = rg.H[i];
  }
}
__device__ __forceinline__ void sb_load_rigid(const float* bc, Rigid& rg) {
#pragma unroll
  for (int i = 0; i < 3; ++i) rg.c[i] = bc[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) rg.R[i] = bc[3 + i];
#pragma unroll
  for (int i = 0; i < 9; ++i) rg.H[i] = bc[12 + i];
}

// role G, whole kernel: for every frame of this CTA, moments partials -> (c, R, H)
__device__ __forceinline__ void sb_geometry_role(SbSmem& s, int n_align, long long nframes, int lane) {
  for (long long it = 0; it < nframes; ++it) {
    const int k = (int)(it & 1);
    mbar_wait_hint(&s.mom_full[k], (uint32_t)((it >> 1) & 1));
    Rigid rg;
    sb_finish_rigid(n_align, s.red[k], rg);
    if (it >= 2) mbar_wait_hint(&s.rig_empty[k], (uint32_t)(((it >> 1) - 1) & 1));
    sb_store_rigid(rg, s.bc[k], lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.rig_full[k]);
  }
}

// Gradient accumulation into the zero-filled dense row with VECTOR reductions (REDG.E.ADD.F32x4 / x2): an atom's three
// floats sit at a 4-byte-aligned address, so depending on its phase inside the 16-byte grid they go out as one 16-byte
// reduction padded with +0.0 (which leaves the neighbouring atom's component unchanged) or as an 8-byte one plus a
// scalar -- 1.5 requests per atom instead of 3.  The scalar REDs were 30 % of the kernel (no-RED probe in profiles/).
struct VecRedAcc {
  float* row;
  int last_atom;                                   // padding must not leave the row
  __device__ __forceinline__ void operator()(int atom, V3 v) {
    float* q = row + 3 * atom;
    const uint32_t ph = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 15u);
    if (ph == 0u && atom != last_atom) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(v.x), "f"(v.y), "f"(v.z), "f"(0.f)
                   : "memory");
    } else if (ph == 4u && atom != 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q - 1), "f"(0.f), "f"(v.x), "f"(v.y), "f"(v.z)
                   : "memory");
    } else if (ph == 8u) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(q), "f"(v.x), "f"(v.y) : "memory");
      atomicAdd(q + 2, v.z);
    } else if (ph == 12u) {
      atomicAdd(q, v.x);
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(q + 1), "f"(v.y), "f"(v.z) : "memory");
    } else {
      atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z);
    }
  }
};

// four consecutive atoms on the 16-byte grid (a backbone dihedral): twelve floats = three 16-byte reductions
__device__ __forceinline__ void acc_quad(VecRedAcc& acc, int a0, int a1, int a2, int a3, V3 g0, V3 g1, V3 g2, V3 g3) {
  float* q = acc.row + 3 * a0;
  if (a1 == a0 + 1 && a2 == a0 + 2 && a3 == a0 + 3 && (reinterpret_cast<uintptr_t>(q) & 15u) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(g0.x), "f"(g0.y), "f"(g0.z), "f"(g1.x)
                 : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 4), "f"(g1.y), "f"(g1.z), "f"(g2.x),
                 "f"(g2.y)
                 : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 8), "f"(g2.z), "f"(g3.x), "f"(g3.y),
                 "f"(g3.z)
                 : "memory");
  } else {
    acc(a0, g0); acc(a1, g1); acc(a2, g2); acc(a3, g3);
  }
}

// feat[L, d] = features(align(x)), one frame per CTA step
__global__ void __launch_bounds__(SB_THREADS, 3)
preprocess_forward_block_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ SbLayout lay,
                                const float* __restrict__ x, float* __restrict__ feat, long long L) {
  extern __shared__ __align__(128) unsigned char smem[];
  SbSmem& s = *reinterpret_cast<SbSmem*>(smem);
  unsigned char* ring = smem + lay.ring_off;
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n3 = 3 * p.n_inp;
  const int stages = lay.stages, buf_bytes = lay.buf_bytes;
  const bool aligned = p.n_align > 0;
  sb_stage_consts(p, lay, smem, s, tid);
  const int n_lead = s.n_lead;
  const long long stride = gridDim.x;
  const long long nframes = (long long)blockIdx.x < L ? (L - 1 - blockIdx.x) / stride + 1 : 0;
  if (warp == 0) {
    if (aligned) sb_geometry_role(s, p.n_align, nframes, lane);
    return;
  }
  // ---- role E ----
  const int et = tid - 32, ewarp = warp - 1;
  auto issue = [&](long long it) {                 // thread 32 only; frame `it` of this CTA -> slot it % stages
    const long long f = blockIdx.x + it * stride;
    if (it < nframes && sb_row_is_bulk(x, f, L, n3)) {
      const int slot = (int)(it % stages);
      const uint32_t bytes = sb_row_bytes(x, f, n3);
      const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(x + f * n3) & 15u);
      mbar_expect_tx(&s.x_full[slot], bytes);
      bulk_g2s(ring + (size_t)slot * buf_bytes, reinterpret_cast<const unsigned char*>(x + f * n3) - off, bytes,
               &s.x_full[slot]);
    }
  };
  uint32_t phases = 0;                            // bit i = parity of stage i
  auto land = [&](long long it) -> const float* {  // wait for (or fetch) frame `it`; returns its row
    const long long f = blockIdx.x + it * stride;
    const int slot = (int)(it % stages);
    unsigned char* buf = ring + (size_t)slot * buf_bytes;
    const float* xf = sb_row_ptr(x, f, n3, buf);
    if (sb_row_is_bulk(x, f, L, n3)) {
      mbar_wait_hint(&s.x_full[slot], (phases >> slot) & 1u);
      phases ^= 1u << slot;
    } else {
      float* dst = const_cast<float*>(xf);
      for (int i = et; i < n3; i += SB_E_THREADS) dst[i] = x[f * n3 + i];
      sb_e_sync();
    }
    return xf;
  };
  if (tid == 32)
    for (int i = 0; i < stages; ++i) issue(i);
  const float* xf = nframes > 0 ? land(0) : nullptr;
  if (aligned && nframes > 0) {
    sb_moments_partial(xf, aidx, ref, p.n_align, et, lane, ewarp, s.red[0]);
    sb_e_sync();
    if (tid == 32) mbar_arrive(&s.mom_full[0]);
  }
  for (long long it = 0; it < nframes; ++it) {
    const long long f = blockIdx.x + it * stride;
    const int k = (int)(it & 1);
    const float* xn = nullptr;
    if (it + 1 < nframes) {
      xn = land(it + 1);
      if (aligned) {
        sb_moments_partial(xn, aidx, ref, p.n_align, et, lane, ewarp, s.red[k ^ 1]);
        sb_e_sync();
        if (tid == 32) mbar_arrive(&s.mom_full[k ^ 1]);
      }
    }
    GlobalOut out{feat + f * p.d_feat};
    Rigid rg;
    if (aligned) {
      for (int e = n_lead + et; e < p.n_entries; e += SB_E_THREADS) {      // entries that do not need the rotation
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        if (en.type != FEAT_POSITION) feature_forward(en, xf, true, rg, p.use_angle, out);
      }
      mbar_wait_hint(&s.rig_full[k], (uint32_t)((it >> 1) & 1));
      sb_load_rigid(s.bc[k], rg);
      for (int e = et; e < p.n_entries; e += SB_E_THREADS) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        if (en.type == FEAT_POSITION) feature_forward(en, xf, true, rg, p.use_angle, out);
      }
    } else {
      for (int e = et; e < p.n_entries; e += SB_E_THREADS) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, false, rg, p.use_angle, out);
      }
    }
    sb_e_sync();                                   // slot and bc[k] consumed by every E thread
    if (tid == 32) {
      if (aligned) mbar_arrive(&s.rig_empty[k]);
      issue(it + stages);
    }
    xf = xn;
  }
}

// gx[L, n, 3] = d<gfeat, features(align(x))>/dx, one frame per CTA step
__global__ void __launch_bounds__(SB_THREADS, 3)
preprocess_backward_block_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ SbLayout lay,
                                 const float* __restrict__ x, const float* __restrict__ gfeat,
                                 float* __restrict__ gx, long long L) {
  extern __shared__ __align__(128) unsigned char smem[];
  SbSmem& s = *reinterpret_cast<SbSmem*>(smem);
  unsigned char* ring = smem + lay.ring_off;
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  const int stages = lay.stages, buf_bytes = lay.buf_bytes;
  const int slot_bytes = lay.buf_bytes + lay.fbuf_bytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n3 = 3 * p.n_inp;
  const bool aligned = p.n_align > 0;
  if (lay.zero_off >= 0) {
    for (int i = tid; i < SB_ZERO_BYTES / 16; i += SB_THREADS)
      reinterpret_cast<uint4*>(smem + lay.zero_off)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  sb_stage_consts(p, lay, smem, s, tid);
  const int n_lead = s.n_lead;
  const long long stride = gridDim.x;
  const long long nframes = (long long)blockIdx.x < L ? (L - 1 - blockIdx.x) / stride + 1 : 0;
  if (warp == 0) {
    if (aligned) sb_geometry_role(s, p.n_align, nframes, lane);
    return;
  }
  // ---- role E ----
  const int et = tid - 32, ewarp = warp - 1;
  auto row_bulk = [&](long long f) { return sb_row_is_bulk(x, f, L, n3) && sb_row_is_bulk(gfeat, f, L, p.d_feat); };
  auto issue = [&](long long it) {                 // thread 32 only
    const long long f = blockIdx.x + it * stride;
    if (it < nframes && row_bulk(f)) {
      const int slot = (int)(it % stages);
      unsigned char* dst = ring + (size_t)slot * slot_bytes;
      const uint32_t bx = sb_row_bytes(x, f, n3), bf = sb_row_bytes(gfeat, f, p.d_feat);
      const uint32_t ox = (uint32_t)(reinterpret_cast<uintptr_t>(x + f * n3) & 15u);
      const uint32_t of = (uint32_t)(reinterpret_cast<uintptr_t>(gfeat + f * p.d_feat) & 15u);
      mbar_expect_tx(&s.x_full[slot], bx + bf);
      bulk_g2s(dst, reinterpret_cast<const unsigned char*>(x + f * n3) - ox, bx, &s.x_full[slot]);
      bulk_g2s(dst + buf_bytes, reinterpret_cast<const unsigned char*>(gfeat + f * p.d_feat) - of, bf,
               &s.x_full[slot]);
    }
  };
  uint32_t phases = 0;
  auto land = [&](long long it) -> const float* {
    const long long f = blockIdx.x + it * stride;
    const int slot = (int)(it % stages);
    unsigned char* buf = ring + (size_t)slot * slot_bytes;
    const float* xf = sb_row_ptr(x, f, n3, buf);
    if (row_bulk(f)) {
      mbar_wait_hint(&s.x_full[slot], (phases >> slot) & 1u);
      phases ^= 1u << slot;
    } else {
      float* dx_ = const_cast<float*>(xf);
      float* df_ = const_cast<float*>(sb_row_ptr(gfeat, f, p.d_feat, buf + buf_bytes));
      for (int i = et; i < n3; i += SB_E_THREADS) dx_[i] = x[f * n3 + i];
      for (int i = et; i < p.d_feat; i += SB_E_THREADS) df_[i] = gfeat[f * p.d_feat + i];
      sb_e_sync();
    }
    return xf;
  };
  if (tid == 32)
    for (int i = 0; i < stages; ++i) issue(i);
  const float* xf = nframes > 0 ? land(0) : nullptr;
  if (aligned && nframes > 0) {
    sb_moments_partial(xf, aidx, ref, p.n_align, et, lane, ewarp, s.red[0]);
    sb_e_sync();
    if (tid == 32) mbar_arrive(&s.mom_full[0]);
  }
  for (long long it = 0; it < nframes; ++it) {
    SB_EVT((int)it, 0);
    const long long f = blockIdx.x + it * stride;
    const int k = (int)(it & 1);
    // dense row of zeros first: plain stores, ordered before this frame's REDs by the role barrier below
    float* dstg = gx + f * n3;
    {
      const int head0 = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dstg) & 15u)) & 15u) >> 2);
      const int head = head0 < n3 ? head0 : n3;
      const int nv = (n3 - head) >> 2;
      if (et < head) dstg[et] = 0.f;
      for (int i = head + 4 * nv + et; i < n3; i += SB_E_THREADS) dstg[i] = 0.f;
      if (lay.zero_off >= 0) {
        // the aligned middle of the row leaves as a few bulk stores of a block of zeros issued by ONE thread: 224
        // threads x 6 STG.128 per frame were competing with the REDs for the LSU
        if (tid == 32) {
          const unsigned char* zsrc = smem + lay.zero_off;
          unsigned char* d = reinterpret_cast<unsigned char*>(dstg + head);
          long long left = 16LL * nv;
          while (left > 0) {
            const uint32_t nb = left < SB_ZERO_BYTES ? (uint32_t)left : (uint32_t)SB_ZERO_BYTES;
            bulk_s2g(d, zsrc, nb);
            d += nb;
            left -= nb;
          }
          bulk_commit();
        }
      } else {
        float4* d4 = reinterpret_cast<float4*>(dstg + head);
        for (int i = et; i < nv; i += SB_E_THREADS) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    SB_EVT((int)it, 1);
    const float* xn = nullptr;
    if (it + 1 < nframes) {
      xn = land(it + 1);
      if (aligned) sb_moments_partial(xn, aidx, ref, p.n_align, et, lane, ewarp, s.red[k ^ 1]);
    }
    if (lay.zero_off >= 0 && tid == 32) bulk_wait0();             // the zeros have landed (not just been read)
    sb_e_sync();                                   // zeros before REDs; moment partials complete
    if (aligned && it + 1 < nframes && tid == 32) mbar_arrive(&s.mom_full[k ^ 1]);
    SB_EVT((int)it, 2);
    const float* gf = sb_row_ptr(gfeat, f, p.d_feat, ring + (size_t)(it % stages) * slot_bytes + buf_bytes);
    SmemGIn gin{gf};
    VecRedAcc acc{dstg, p.n_inp - 1};
    Rigid rg;
    float v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = 0.f;
    float (&M)[9] = *reinterpret_cast<float (*)[9]>(&v[0]);
    float (&sg)[3] = *reinterpret_cast<float (*)[3]>(&v[9]);
    if (aligned) {
      for (int e = n_lead + et; e < p.n_entries; e += SB_E_THREADS) {      // entries that do not need the rotation
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        if (en.type != FEAT_POSITION) feature_backward(en, xf, true, rg, p.use_angle, gin, acc, M, sg);
      }
      SB_EVT((int)it, 3);
      mbar_wait_hint(&s.rig_full[k], (uint32_t)((it >> 1) & 1));
      SB_EVT((int)it, 4);
      sb_load_rigid(s.bc[k], rg);
      for (int e = et; e < p.n_entries; e += SB_E_THREADS) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        if (en.type == FEAT_POSITION) feature_backward(en, xf, true, rg, p.use_angle, gin, acc, M, sg);
      }
      sb_reduce12_store(v, s.redm + ewarp * 12, lane);
      SB_EVT((int)it, 5);
      sb_e_sync();                                 // also: every read of the x / cotangent slot is done
      SB_EVT((int)it, 6);
      if (tid == 32) issue(it + stages);
      if (ewarp == 0) {
        float Ms[9], sgs[3];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < SB_E_WARPS; ++w) a += s.redm[w * 12 + i];
          if (i < 9) Ms[i] = a; else sgs[i - 9] = a;
        }
        float dH[9];
        align_backward_dH(rg, Ms, dH);
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 9; ++i) s.bcm[i] = dH[i];
#pragma unroll
          for (int i = 0; i < 3; ++i) s.bcm[9 + i] = sgs[i];
        }
      }
      SB_EVT((int)it, 7);
      sb_e_sync();
      SB_EVT((int)it, 8);
      if (tid == 32) mbar_arrive(&s.rig_empty[k]);               // every E thread holds (c, R, H) in registers
      float dH[9], sgs[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) dH[i] = s.bcm[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) sgs[i] = s.bcm[9 + i];
      const float inv_na = 1.0f / (float)p.n_align;
      for (int kk = et; kk < p.n_align; kk += SB_E_THREADS)
        acc(aidx[kk], align_atom_grad(dH, sgs, inv_na, ref[3 * kk], ref[3 * kk + 1], ref[3 * kk + 2]));
      SB_EVT((int)it, 9);
      sb_e_sync();                                 // redm / bcm reusable by the next frame
    } else {
      for (int e = et; e < p.n_entries; e += SB_E_THREADS) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_backward(en, xf, false, rg, p.use_angle, gin, acc, M, sg);
      }
      sb_e_sync();
      if (tid == 32) issue(it + stages);
    }
    SB_EVT((int)it, 10);
    xf = xn;
  }
}

}  // namespace molann
