// general.cuh -- size-agnostic kernels: warp-per-frame geometry straight from global memory and
// layered FFMA GEMMs for MLPs whose activations do not fit on chip next to a frame tile.
//
// Geometry: 32 lanes stride over alignment atoms / feature entries; centroid, covariance and the
// backward moment are warp-shuffle (xor butterfly) reductions; the coordinate gradient is scattered
// with fire-and-forget REDs into a row the same warp has just zero-filled (both land in L2, so DRAM
// sees one write of the dense [n,3] row).
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace molann {

constexpr int WARPS_PER_CTA = 8;

struct GlobalOut {
  float* row;
  __device__ __forceinline__ void operator()(int col, float v) { row[col] = v; }
};
struct GlobalGIn {
  const float* row;
  __device__ __forceinline__ float operator()(int col) const { return __ldg(row + col); }
  __device__ __forceinline__ void load2(int col, float& a, float& b) const { a = (*this)(col); b = (*this)(col + 1); }
  __device__ __forceinline__ void load3(int col, float& a, float& b, float& c) const {
    a = (*this)(col); b = (*this)(col + 1); c = (*this)(col + 2);
  }
};
struct RedAcc {
  float* row;
  __device__ __forceinline__ void operator()(int atom, V3 v) {
    float* q = row + 3 * atom;
#ifdef MOLANN_PROBE_NO_RED                      // tests/cuda/sb_trace.cu: what the REDs cost (results are wrong)
    if (v.x == 123456.f) q[0] = v.y + v.z;
#else
    atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z);
#endif
  }
};

// feat[L, d] = features(align(x))
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
preprocess_forward_warp_kernel(const __grid_constant__ DevPlan p, const float* __restrict__ x,
                               float* __restrict__ feat, long long L) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * WARPS_PER_CTA;
  const int n3 = 3 * p.n_inp;
  const bool aligned = p.n_align > 0;
  for (long long f = warp0; f < L; f += nwarps) {
    const float* xf = x + f * n3;
    Rigid rg;
    if (aligned) kabsch<32>(xf, p.align_idx, p.ref_x, p.n_align, lane, rg);
    GlobalOut out{feat + f * p.d_feat};
    for (int e = lane; e < p.n_entries; e += 32) {
      const Entry en = load_entry(p.entries + ENTRY_INTS * e);
      feature_forward(en, xf, aligned, rg, p.use_angle, out);
    }
  }
}

// gx[L, n, 3] = d<gfeat, features(align(x))>/dx
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
preprocess_backward_warp_kernel(const __grid_constant__ DevPlan p, const float* __restrict__ x,
                                const float* __restrict__ gfeat, float* __restrict__ gx, long long L) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * WARPS_PER_CTA;
  const int n3 = 3 * p.n_inp;
  const bool aligned = p.n_align > 0;
  for (long long f = warp0; f < L; f += nwarps) {
    const float* xf = x + f * n3;
    float* gxf = gx + f * n3;
    for (int i = lane; i < n3; i += 32) gxf[i] = 0.f;
    __syncwarp();
    Rigid rg;
    if (aligned) kabsch<32>(xf, p.align_idx, p.ref_x, p.n_align, lane, rg);
    GlobalGIn gin{gfeat + f * p.d_feat};
    RedAcc acc{gxf};
    float M[9], sg[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.f;
    sg[0] = sg[1] = sg[2] = 0.f;
    for (int e = lane; e < p.n_entries; e += 32) {
      const Entry en = load_entry(p.entries + ENTRY_INTS * e);
      feature_backward(en, xf, aligned, rg, p.use_angle, gin, acc, M, sg);
    }
    if (aligned) {
#pragma unroll
      for (int i = 0; i < 9; ++i) M[i] = gsum<32>(M[i]);
#pragma unroll
      for (int i = 0; i < 3; ++i) sg[i] = gsum<32>(sg[i]);
      float dH[9];
      align_backward_dH(rg, M, dH);
      const float inv_na = 1.0f / (float)p.n_align;
      for (int k = lane; k < p.n_align; k += 32)
        acc(p.align_idx[k],
            align_atom_grad(dH, sg, inv_na, p.ref_x[3 * k], p.ref_x[3 * k + 1], p.ref_x[3 * k + 2]));
    }
  }
}

// out[L, n, 3] = (x - c) R   (stand-alone AlignmentLayer.forward)
// ---------------------------------------------------------------------------------------------------------
// Staged variants for big systems (C3 / C5).  The gather kernels above read ~300 scattered atoms of a 24 KB frame
// twice (moments, then features); DRAM fetches 64 bytes per touched 12-byte atom, so they move 2x the frame
// (ncu: 1.57 GB for a 786 MB batch, 88 % of the HBM peak -- memory bound on bytes nobody needs).  Here every warp
// owns a frame-sized shared-memory buffer, pulls the frame in ONCE with one contiguous bulk copy
// (cp.async.bulk from the 16-byte boundary below the frame) and gathers from shared memory.  The backward also
// builds the dense gradient row in shared memory (shared-memory atomics) and writes it out with one bulk store,
// instead of zero-filling the row in HBM and merging scattered REDs in L2.
// ---------------------------------------------------------------------------------------------------------
struct SmemRedAcc {
  float* row;
  __device__ __forceinline__ void operator()(int atom, V3 v) {
    float* q = row + 3 * atom;
    atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z);
  }
};

struct SmemGIn {
  const float* row;
  __device__ __forceinline__ float operator()(int col) const { return row[col]; }
  __device__ __forceinline__ void load2(int col, float& a, float& b) const { a = row[col]; b = row[col + 1]; }
  __device__ __forceinline__ void load3(int col, float& a, float& b, float& c) const {
    a = row[col]; b = row[col + 1]; c = row[col + 2];
  }
};

// one warp: row f (n3 floats) of x -> its buffer; returns the row's address inside the buffer
__device__ __forceinline__ const float* stage_frame(const float* __restrict__ x, long long f, long long L, int n3,
                                                    unsigned char* buf, unsigned long long* bar, uint32_t& phase,
                                                    int lane) {
  const float* src = x + f * n3;
  const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
  float* dst = reinterpret_cast<float*>(buf + off);
  if (off == 0u || f + 1 < L) {               // the copy may run up to 15 bytes past the frame: not on the last one
    if (lane == 0) {
      const uint32_t bytes = ((uint32_t)n3 * 4u + off + 15u) & ~15u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(buf, reinterpret_cast<const unsigned char*>(src) - off, bytes, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
  } else {
    for (int i = lane; i < n3; i += 32) dst[i] = src[i];
    __syncwarp();
  }
  return dst;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
preprocess_forward_staged_kernel(const __grid_constant__ DevPlan p, const float* __restrict__ x,
                                 float* __restrict__ feat, long long L, int buf_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
  unsigned char* buf = smem + 128 + (size_t)w * buf_bytes;
  if (lane == 0) {
    mbar_init(&bars[w], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int n3 = 3 * p.n_inp;
  const bool aligned = p.n_align > 0;
  uint32_t phase = 0;
  for (long long f = (long long)blockIdx.x * nw + w; f < L; f += (long long)gridDim.x * nw) {
    const float* xf = stage_frame(x, f, L, n3, buf, &bars[w], phase, lane);
    Rigid rg;
    if (aligned) kabsch<32>(xf, p.align_idx, p.ref_x, p.n_align, lane, rg);
    GlobalOut out{feat + f * p.d_feat};
    for (int e = lane; e < p.n_entries; e += 32) {
      const Entry en = load_entry(p.entries + ENTRY_INTS * e);
      feature_forward(en, xf, aligned, rg, p.use_angle, out);
    }
    __syncwarp();                              // every lane is done with the buffer before the next copy lands
  }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
preprocess_backward_staged_kernel(const __grid_constant__ DevPlan p, const float* __restrict__ x,
                                  const float* __restrict__ gfeat, float* __restrict__ gx, long long L, int buf_bytes,
                                  int fbuf_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
  unsigned char* buf = smem + 128 + (size_t)w * (2 * buf_bytes + fbuf_bytes);   // [x frame | gradient row | cotangent]
  unsigned char* gbuf = buf + buf_bytes;
  unsigned char* fbuf = gbuf + buf_bytes;
  if (lane == 0) {
    mbar_init(&bars[2 * w], 1);
    mbar_init(&bars[2 * w + 1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int n3 = 3 * p.n_inp;
  const bool aligned = p.n_align > 0;
  uint32_t phase = 0, phase_f = 0;
  for (long long f = (long long)blockIdx.x * nw + w; f < L; f += (long long)gridDim.x * nw) {
    const float* xf = stage_frame(x, f, L, n3, buf, &bars[2 * w], phase, lane);
    const float* gf = stage_frame(gfeat, f, L, p.d_feat, fbuf, &bars[2 * w + 1], phase_f, lane);
    float* dstg = gx + f * n3;
    const uint32_t goff = (uint32_t)(reinterpret_cast<uintptr_t>(dstg) & 15u);
    float* grow = reinterpret_cast<float*>(gbuf + goff);            // same 16-byte phase as the destination
    if (lane == 0) bulk_wait_read0();          // the previous row's bulk store has finished reading the buffer
    __syncwarp();
    for (int i = lane; i < buf_bytes / 16; i += 32) reinterpret_cast<uint4*>(gbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    Rigid rg;
    if (aligned) kabsch<32>(xf, p.align_idx, p.ref_x, p.n_align, lane, rg);
    SmemGIn gin{gf};
    SmemRedAcc acc{grow};
    float M[9], sg[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.f;
    sg[0] = sg[1] = sg[2] = 0.f;
    for (int e = lane; e < p.n_entries; e += 32) {
      const Entry en = load_entry(p.entries + ENTRY_INTS * e);
      feature_backward(en, xf, aligned, rg, p.use_angle, gin, acc, M, sg);
    }
    if (aligned) {
#pragma unroll
      for (int i = 0; i < 9; ++i) M[i] = gsum<32>(M[i]);
#pragma unroll
      for (int i = 0; i < 3; ++i) sg[i] = gsum<32>(sg[i]);
      float dH[9];
      align_backward_dH(rg, M, dH);
      const float inv_na = 1.0f / (float)p.n_align;
      for (int k = lane; k < p.n_align; k += 32)
        acc(p.align_idx[k],
            align_atom_grad(dH, sg, inv_na, p.ref_x[3 * k], p.ref_x[3 * k + 1], p.ref_x[3 * k + 2]));
    }
    __syncwarp();
    // write the row: head / tail up to the 16-byte grid with plain stores, the aligned middle as one bulk store
    const int head = goff ? (int)((16u - goff) >> 2) : 0;            // floats before the first 16-byte boundary
    const int mid = ((n3 - head) >> 2) << 2;
    for (int i = lane; i < head && i < n3; i += 32) dstg[i] = grow[i];
    for (int i = head + mid + lane; i < n3; i += 32) dstg[i] = grow[i];
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && mid > 0) {
      bulk_s2g(dstg + head, grow + head, (uint32_t)mid * 4u);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait0();
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
align_forward_warp_kernel(const __grid_constant__ DevPlan p, const float* __restrict__ x, float* __restrict__ out,
                          long long L) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * WARPS_PER_CTA;
  const int n3 = 3 * p.n_inp;
  for (long long f = warp0; f < L; f += nwarps) {
    const float* xf = x + f * n3;
    float* of = out + f * n3;
    Rigid rg;
    kabsch<32>(xf, p.align_idx, p.ref_x, p.n_align, lane, rg);
    for (int j = lane; j < p.n_inp; j += 32) {
      float zx, zy, zz;
      rigid_apply(rg, xf[3 * j], xf[3 * j + 1], xf[3 * j + 2], zx, zy, zz);
      of[3 * j] = zx; of[3 * j + 1] = zy; of[3 * j + 2] = zz;
    }
  }
}

// gx = d<gout, align(x)>/dx
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
align_backward_warp_kernel(const __grid_constant__ DevPlan p, const float* __restrict__ x,
                           const float* __restrict__ gout, float* __restrict__ gx, long long L) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * WARPS_PER_CTA;
  const int n3 = 3 * p.n_inp;
  for (long long f = warp0; f < L; f += nwarps) {
    const float* xf = x + f * n3;
    const float* gf = gout + f * n3;
    float* gxf = gx + f * n3;
    Rigid rg;
    kabsch<32>(xf, p.align_idx, p.ref_x, p.n_align, lane, rg);
    float M[9], sg[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.f;
    sg[0] = sg[1] = sg[2] = 0.f;
    for (int j = lane; j < p.n_inp; j += 32) {
      const float g0 = __ldg(gf + 3 * j), g1 = __ldg(gf + 3 * j + 1), g2 = __ldg(gf + 3 * j + 2);
      const float dx = xf[3 * j] - rg.c[0], dy = xf[3 * j + 1] - rg.c[1], dz = xf[3 * j + 2] - rg.c[2];
      M[0] = fmaf(dx, g0, M[0]); M[1] = fmaf(dx, g1, M[1]); M[2] = fmaf(dx, g2, M[2]);
      M[3] = fmaf(dy, g0, M[3]); M[4] = fmaf(dy, g1, M[4]); M[5] = fmaf(dy, g2, M[5]);
      M[6] = fmaf(dz, g0, M[6]); M[7] = fmaf(dz, g1, M[7]); M[8] = fmaf(dz, g2, M[8]);
      float tx, ty, tz;
      rot_transpose_apply(rg, g0, g1, g2, tx, ty, tz);
      sg[0] += tx; sg[1] += ty; sg[2] += tz;
      gxf[3 * j] = tx; gxf[3 * j + 1] = ty; gxf[3 * j + 2] = tz;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = gsum<32>(M[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) sg[i] = gsum<32>(sg[i]);
    float dH[9];
    align_backward_dH(rg, M, dH);
    const float inv_na = 1.0f / (float)p.n_align;
    RedAcc acc{gxf};
    for (int k = lane; k < p.n_align; k += 32)
      acc(p.align_idx[k], align_atom_grad(dH, sg, inv_na, p.ref_x[3 * k], p.ref_x[3 * k + 1], p.ref_x[3 * k + 2]));
  }
}

// ---------------------------------------------------------------------------------------------
// Tiled FFMA GEMM, C[M,N] = epilogue( sum_k A(m,k) * B(k,n) ), 64x64x16 tiles, 4x4 per thread.
//   A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]  (generic strides cover NT / NN / TN).
// Epilogues:
//   EPI_BIAS_ACT : C = act(acc + bias[n])                        (Linear forward)
//   EPI_DACT     : C = acc * act'(Hprev[m,n])   (Hprev may be null: plain store)   (backward to input)
//   EPI_ATOMIC   : atomicAdd(C, acc)            (split-K over blockIdx.z)          (weight gradient)
// ---------------------------------------------------------------------------------------------
enum { EPI_BIAS_ACT = 0, EPI_DACT = 1, EPI_ATOMIC = 2 };

template <int EPI>
__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
            long long sbn, float* __restrict__ C, long long ldc, int M, int N, long long K, long long k_chunk,
            const float* __restrict__ bias, const float* __restrict__ Hprev, int act, int apply_act) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const long long k_begin = (long long)blockIdx.z * k_chunk;
  const long long k_end = (k_begin + k_chunk < K) ? (k_begin + k_chunk) : K;
  const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, each 4 (m) x 4 (n)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: prefer the unit-stride direction of each operand for coalescing
  const bool a_k_contig = (sak == 1);
  const bool b_n_contig = (sbn == 1);
  for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      int kk, mm;
      if (a_k_contig) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const long long gk = k0 + kk;
      const int gm = m0 + mm;
      As[kk][mm] = (gm < M && gk < k_end) ? __ldg(A + gm * sam + gk * sak) : 0.f;
      int kb, nn;
      if (b_n_contig) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
      const long long gkb = k0 + kb;
      const int gn = n0 + nn;
      Bs[kb][nn] = (gn < N && gkb < k_end) ? __ldg(B + gkb * sbk + gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float* c = C + (long long)gm * ldc + gn;
      if (EPI == EPI_BIAS_ACT) {
        v += __ldg(bias + gn);
        if (apply_act) v = act_forward(v, act);
        *c = v;
      } else if (EPI == EPI_DACT) {
        if (apply_act) v *= act_grad_from_output(__ldg(Hprev + (long long)gm * ldc + gn), act);
        *c = v;
      } else {
        atomicAdd(c, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Narrow layers (the model's last Linear: 128 -> 2 for every BASELINE config).  The 64 x 64 tile kernel spends a
// whole tile on 2 output columns (33 us per 32768 x 128 -> 2, 22 us for the matching backward); these two stream the
// wide operand once at full width instead.
// ---------------------------------------------------------------------------------------------
constexpr int NARROW_MAX = 8;

// out[m, o] = act(b[o] + sum_k in[m, k] W[o, k]),  N <= NARROW_MAX: one warp per row, lanes stride k
__global__ void __launch_bounds__(256)
narrow_forward_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ b,
                      float* __restrict__ out, long long M, int K, int N, int act, int apply_act) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  const bool vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(W) & 15u) == 0);
  for (long long m = warp0; m < M; m += nwarps) {
    const float* row = in + m * K;
    float acc[NARROW_MAX];
#pragma unroll
    for (int o = 0; o < NARROW_MAX; ++o) acc[o] = 0.f;
    if (vec) {
      for (int k = 4 * lane; k < K; k += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
#pragma unroll
        for (int o = 0; o < NARROW_MAX; ++o) {
          if (o < N) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)o * K + k));
            acc[o] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[o]))));
          }
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float v = __ldg(row + k);
#pragma unroll
        for (int o = 0; o < NARROW_MAX; ++o)
          if (o < N) acc[o] = fmaf(v, __ldg(W + (long long)o * K + k), acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < NARROW_MAX; ++o)
      if (o < N) acc[o] = gsum<32>(acc[o]);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < NARROW_MAX; ++o) {
        if (o < N) {
          float v = acc[o] + __ldg(b + o);
          if (apply_act) v = act_forward(v, act);
          out[m * N + o] = v;
        }
      }
    }
  }
}

// gprev[m, k] = (sum_o gz[m, o] W[o, k]) * act'(hprev[m, k]),  N <= NARROW_MAX: one thread per four consecutive k
__global__ void __launch_bounds__(256)
narrow_backward_input_kernel(const float* __restrict__ gz, const float* __restrict__ W,
                             const float* __restrict__ hprev, float* __restrict__ gprev, long long M, int K, int N,
                             int act) {
  const bool vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(gprev) & 15u) == 0) &&
                   (hprev == nullptr || (reinterpret_cast<uintptr_t>(hprev) & 15u) == 0);
  if (vec) {
    const int K4 = K >> 2;
    const long long total = M * K4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const long long m = i / K4;
      const int k = (int)(i - m * K4) << 2;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NARROW_MAX; ++o) {
        if (o < N) {
          const float g = __ldg(gz + m * N + o);
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)o * K + k));
          a.x = fmaf(g, w.x, a.x); a.y = fmaf(g, w.y, a.y); a.z = fmaf(g, w.z, a.z); a.w = fmaf(g, w.w, a.w);
        }
      }
      if (hprev != nullptr) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(hprev + m * K + k));
        a.x *= act_grad_from_output(h.x, act); a.y *= act_grad_from_output(h.y, act);
        a.z *= act_grad_from_output(h.z, act); a.w *= act_grad_from_output(h.w, act);
      }
      *reinterpret_cast<float4*>(gprev + m * K + k) = a;
    }
    return;
  }
  const long long total = M * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / K;
    const int k = (int)(i - m * K);
    float a = 0.f;
#pragma unroll
    for (int o = 0; o < NARROW_MAX; ++o)
      if (o < N) a = fmaf(__ldg(gz + m * N + o), __ldg(W + (long long)o * K + k), a);
    if (hprev != nullptr) a *= act_grad_from_output(__ldg(hprev + i), act);
    gprev[i] = a;
  }
}

// gb[n] += sum_m gz[m, n]
__global__ void __launch_bounds__(256)
colsum_atomic_kernel(const float* __restrict__ gz, int M, int N, float* __restrict__ gb, int rows_per_block) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = r0 + (threadIdx.x >> 5); m < r1; m += 8) s += __ldg(gz + (long long)m * N + n);
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    if (n < N) atomicAdd(gb + n, t);
  }
}

}  // namespace molann
