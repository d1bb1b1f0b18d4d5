// small_tile.cuh -- stand-alone preprocessing (PreprocessingANN.forward / backward, molann/ann.py:553-565) for SMALL
// frames: thread = frame over a shared-memory tile.
//
// The warp-per-frame kernels of general.cuh spend a whole warp on a 22-atom frame: 32 lanes for 10 alignment atoms
// and 10 entries, five shuffle rounds per reduced value -- 290 us per 196608 C2 frames (180 GB/s; it was the largest
// kernel of a C4 training step, which needs the features as reconstruction target and as layer-0 activation).  Here
// a CTA stages F frames (one contiguous byte range: one bulk copy, or coalesced loads when the range is not 16-byte
// aligned) and every thread walks ONE frame: moments, rotation and feature program with no cross-lane traffic, the
// same device functions the fused kernels use.  Results leave through a shared-memory tile so global stores are
// full rows of contiguous bytes.
#pragma once
#include "common.cuh"
#include "general.cuh"
#include "geometry.cuh"

namespace molann {

constexpr int ST_F = 128;                          // frames per tile = threads per CTA

struct StLayout {                                  // byte offsets into dynamic shared memory (host-computed)
  int aidx_off, ref_off, ent_off;
  int xs_off;                                      // [F][3n] coordinates
  int out_off;                                     // forward: [F][d_feat] features; backward: [F][3n] gradient rows
  int gf_off;                                      // backward: [F][d_feat] cotangent rows
  int total;
};

struct StRowOut {
  float* row;
  __device__ __forceinline__ void operator()(int col, float v) { row[col] = v; }
};
struct StRowAcc {                                  // the row belongs to this thread alone: plain read-modify-write
  float* row;
  __device__ __forceinline__ void operator()(int atom, V3 v) {
    float* q = row + 3 * atom;
    q[0] += v.x; q[1] += v.y; q[2] += v.z;
  }
};

__device__ __forceinline__ void st_stage_consts(const DevPlan& p, const StLayout& lay, unsigned char* smem, int tid) {
  int* aidx = reinterpret_cast<int*>(smem + lay.aidx_off);
  float* ref = reinterpret_cast<float*>(smem + lay.ref_off);
  int* ent = reinterpret_cast<int*>(smem + lay.ent_off);
  for (int i = tid; i < p.n_align; i += ST_F) aidx[i] = __ldg(p.align_idx + i);
  for (int i = tid; i < 3 * p.n_align; i += ST_F) ref[i] = __ldg(p.ref_x + i);
  for (int i = tid; i < ENTRY_INTS * p.n_entries; i += ST_F) ent[i] = __ldg(p.entries + i);
}

// rows [f0, f0 + nf) of a [L, w] matrix -> dst (contiguous), all threads
__device__ __forceinline__ void st_load_rows(const float* __restrict__ src, long long f0, int nf, int w, float* dst,
                                             int tid) {
  const float* s = src + f0 * w;
  const int n = nf * w;
  if ((reinterpret_cast<uintptr_t>(s) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < (n >> 2); i += ST_F) d4[i] = __ldg(s4 + i);
    for (int i = (n & ~3) + tid; i < n; i += ST_F) dst[i] = __ldg(s + i);
  } else {
    for (int i = tid; i < n; i += ST_F) dst[i] = __ldg(s + i);
  }
}
__device__ __forceinline__ void st_store_rows(float* __restrict__ dstg, long long f0, int nf, int w, const float* src,
                                              int tid) {
  float* d = dstg + f0 * w;
  const int n = nf * w;
  if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int i = tid; i < (n >> 2); i += ST_F) d4[i] = s4[i];
    for (int i = (n & ~3) + tid; i < n; i += ST_F) d[i] = src[i];
  } else {
    for (int i = tid; i < n; i += ST_F) d[i] = src[i];
  }
}

// feat[L, d] = features(align(x))
__global__ void __launch_bounds__(ST_F)
preprocess_forward_tile_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ StLayout lay,
                               const float* __restrict__ x, float* __restrict__ feat, long long L) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  float* outs = reinterpret_cast<float*>(smem + lay.out_off);
  st_stage_consts(p, lay, smem, tid);
  const bool aligned = p.n_align > 0;
  const long long ntiles = (L + ST_F - 1) / ST_F;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long f0 = t * ST_F;
    const int nf = (int)((L - f0) < ST_F ? (L - f0) : ST_F);
    __syncthreads();                               // previous tile fully written out; constants staged
    st_load_rows(x, f0, nf, n3, xs, tid);
    __syncthreads();
    {                                              // every lane runs (the rotation's fallback votes warp-wide);
      const float* xf = xs + (tid < nf ? tid : nf - 1) * n3;       // lanes past the batch redo its last frame
      Rigid rg;
      if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      StRowOut out{outs + tid * p.d_feat};
      for (int e = 0; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    __syncthreads();
    st_store_rows(feat, f0, nf, p.d_feat, outs, tid);
  }
}

// gx[L, n, 3] = d<gfeat, features(align(x))>/dx
__global__ void __launch_bounds__(ST_F)
preprocess_backward_tile_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ StLayout lay,
                                const float* __restrict__ x, const float* __restrict__ gfeat,
                                float* __restrict__ gx, long long L) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  float* gs = reinterpret_cast<float*>(smem + lay.out_off);
  float* gfs = reinterpret_cast<float*>(smem + lay.gf_off);
  st_stage_consts(p, lay, smem, tid);
  const bool aligned = p.n_align > 0;
  const long long ntiles = (L + ST_F - 1) / ST_F;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long f0 = t * ST_F;
    const int nf = (int)((L - f0) < ST_F ? (L - f0) : ST_F);
    __syncthreads();
    st_load_rows(x, f0, nf, n3, xs, tid);
    st_load_rows(gfeat, f0, nf, p.d_feat, gfs, tid);
    for (int i = tid; i < ST_F * n3; i += ST_F) gs[i] = 0.f;
    __syncthreads();
    {                                              // every lane runs; rows past the batch are never stored
      const int fr = tid < nf ? tid : nf - 1;
      const float* xf = xs + fr * n3;
      Rigid rg;
      if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      SmemGIn gin{gfs + fr * p.d_feat};
      StRowAcc acc{gs + tid * n3};
      float M[9], sg[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) M[i] = 0.f;
      sg[0] = sg[1] = sg[2] = 0.f;
      for (int e = 0; e < p.n_entries; ++e) {
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
    __syncthreads();
    st_store_rows(gx, f0, nf, n3, gs, tid);
  }
}

// out[L, n, 3] = (x - c) R   (stand-alone AlignmentLayer.forward, molann/ann.py:157-199), thread = frame
__global__ void __launch_bounds__(ST_F)
align_forward_tile_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ StLayout lay,
                          const float* __restrict__ x, float* __restrict__ out, long long L) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  st_stage_consts(p, lay, smem, tid);
  const long long ntiles = (L + ST_F - 1) / ST_F;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long f0 = t * ST_F;
    const int nf = (int)((L - f0) < ST_F ? (L - f0) : ST_F);
    __syncthreads();
    st_load_rows(x, f0, nf, n3, xs, tid);
    __syncthreads();
    {
      float* xf = xs + (tid < nf ? tid : nf - 1) * n3;
      Rigid rg;
      kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      if (tid < nf)                                // in place: the row belongs to this thread
        for (int j = 0; j < p.n_inp; ++j) {
          float zx, zy, zz;
          rigid_apply(rg, xf[3 * j], xf[3 * j + 1], xf[3 * j + 2], zx, zy, zz);
          xf[3 * j] = zx; xf[3 * j + 1] = zy; xf[3 * j + 2] = zz;
        }
    }
    __syncthreads();
    st_store_rows(out, f0, nf, n3, xs, tid);
  }
}

// gx = d<gout, align(x)>/dx, thread = frame (closed form)
__global__ void __launch_bounds__(ST_F)
align_backward_tile_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ StLayout lay,
                           const float* __restrict__ x, const float* __restrict__ gout, float* __restrict__ gx,
                           long long L) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  float* gs = reinterpret_cast<float*>(smem + lay.out_off);
  st_stage_consts(p, lay, smem, tid);
  const long long ntiles = (L + ST_F - 1) / ST_F;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long f0 = t * ST_F;
    const int nf = (int)((L - f0) < ST_F ? (L - f0) : ST_F);
    __syncthreads();
    st_load_rows(x, f0, nf, n3, xs, tid);
    st_load_rows(gout, f0, nf, n3, gs, tid);
    __syncthreads();
    {
      const int fr = tid < nf ? tid : nf - 1;
      const float* xf = xs + fr * n3;
      float* gf = gs + fr * n3;
      Rigid rg;
      kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      if (tid < nf) {
        float M[9], sg[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i] = 0.f;
        sg[0] = sg[1] = sg[2] = 0.f;
        for (int j = 0; j < p.n_inp; ++j) {
          const float g0 = gf[3 * j], g1 = gf[3 * j + 1], g2 = gf[3 * j + 2];
          const float dx = xf[3 * j] - rg.c[0], dy = xf[3 * j + 1] - rg.c[1], dz = xf[3 * j + 2] - rg.c[2];
          M[0] = fmaf(dx, g0, M[0]); M[1] = fmaf(dx, g1, M[1]); M[2] = fmaf(dx, g2, M[2]);
          M[3] = fmaf(dy, g0, M[3]); M[4] = fmaf(dy, g1, M[4]); M[5] = fmaf(dy, g2, M[5]);
          M[6] = fmaf(dz, g0, M[6]); M[7] = fmaf(dz, g1, M[7]); M[8] = fmaf(dz, g2, M[8]);
          float tx, ty, tz;
          rot_transpose_apply(rg, g0, g1, g2, tx, ty, tz);
          sg[0] += tx; sg[1] += ty; sg[2] += tz;
          gf[3 * j] = tx; gf[3 * j + 1] = ty; gf[3 * j + 2] = tz;      // in place: G_j -> G_j R^T
        }
        float dH[9];
        align_backward_dH(rg, M, dH);
        const float inv_na = 1.0f / (float)p.n_align;
        StRowAcc acc{gf};
        for (int k = 0; k < p.n_align; ++k)
          acc(aidx[k], align_atom_grad(dH, sg, inv_na, ref[3 * k], ref[3 * k + 1], ref[3 * k + 2]));
      }
    }
    __syncthreads();
    st_store_rows(gx, f0, nf, n3, gs, tid);
  }
}

}  // namespace molann
