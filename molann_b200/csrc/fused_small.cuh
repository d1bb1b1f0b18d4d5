// fused_small.cuh -- one fused kernel per frame tile for small systems (whole tile + MLP fit in smem).
//
//   forward : TMA bulk stage-in of F frames -> thread-per-frame Kabsch/Jacobi + feature program ->
//             register-tiled FFMA MLP with activations resident in shared memory -> y
//   backward: same stage-in, forward recompute (hidden activations kept in smem), MLP backward in
//             place, feature + alignment backward into an smem gradient tile, TMA bulk store of gx.
//
// HBM traffic per frame is exactly the algorithmic 12 n + 4 k (fwd) / 24 n + 8 k (fwd + d/dx) bytes:
// aligned coordinates, features, activations and weights never leave the SM.
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace molann {

// ---------------------------------------------------------------------------------------------
// Register-tiled micro GEMM on smem operands.  A: [K][F] (feature-major activations), B: [K][ldb]
// (n contiguous).  Each work item owns TM frames x 8 outputs.
// ---------------------------------------------------------------------------------------------
template <int TM, int F>
__device__ __forceinline__ void mt_accumulate(const float* __restrict__ A, const float* __restrict__ B, int ldb,
                                              int K, int f0, int n0, float (&acc)[TM][8]) {
  const float* ap = A + f0;
  const float* bp = B + n0;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[TM];
    if constexpr (TM == 8) {
      const float4 a0 = *reinterpret_cast<const float4*>(ap);
      const float4 a1 = *reinterpret_cast<const float4*>(ap + 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    } else if constexpr (TM == 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(ap);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
    } else if constexpr (TM == 2) {
      const float2 a0 = *reinterpret_cast<const float2*>(ap);
      a[0] = a0.x; a[1] = a0.y;
    } else {
      a[0] = ap[0];
    }
    const float4 b0 = *reinterpret_cast<const float4*>(bp);
    const float4 b1 = *reinterpret_cast<const float4*>(bp + 4);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    ap += F;
    bp += ldb;
  }
}

template <int TM, int F>
__device__ __forceinline__ void store_rows(float* __restrict__ dst, const float (&v)[TM]) {
  if constexpr (TM == 8) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else if constexpr (TM == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (TM == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  } else {
    dst[0] = v[0];
  }
}

// out[n][f] = act( sum_k in[k][f] * Wt[k][n] + b[n] );  last layer writes y (global) instead.
template <int TM, int F, int NT>
__device__ __forceinline__ void layer_forward(const float* __restrict__ in, float* __restrict__ out,
                                              const float* __restrict__ Wt, const float* __restrict__ bs, int ldw,
                                              int K, int N, int act, bool last, float* __restrict__ y_tile,
                                              int nf_valid, int tid) {
  constexpr int NFG = F / TM;
  const int NOG = (N + 7) >> 3;
  for (int item = tid; item < NFG * NOG; item += NT) {
    const int fg = item % NFG, og = item / NFG;
    const int f0 = fg * TM, n0 = og * 8;
    float acc[TM][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bj = bs[n0 + j];
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[i][j] = bj;
    }
    mt_accumulate<TM, F>(in, Wt, ldw, K, f0, n0, acc);
    if (last) {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int f = f0 + i;
        if (f < nf_valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j < N) y_tile[(long long)f * N + n0 + j] = acc[i][j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) v[i] = act_forward(acc[i][j], act);
        store_rows<TM, F>(out + (n0 + j) * F + f0, v);
      }
    }
  }
}

// io[i][f] = ( sum_o gz[o][f] * Wn[o][i] ) * act'(io[i][f])      (in place on the layer input)
template <int TM, int F, int NT>
__device__ __forceinline__ void layer_backward(const float* __restrict__ gz, float* __restrict__ io,
                                               const float* __restrict__ Wn, int ldw, int Nout, int Kin, int act,
                                               bool apply_act_grad, int tid) {
  constexpr int NFG = F / TM;
  const int NIG = (Kin + 7) >> 3;
  for (int item = tid; item < NFG * NIG; item += NT) {
    const int fg = item % NFG, ig = item / NFG;
    const int f0 = fg * TM, i0 = ig * 8;
    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    mt_accumulate<TM, F>(gz, Wn, ldw, Nout, f0, i0, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (i0 + j < Kin) {
        float* p = io + (i0 + j) * F + f0;
        float v[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) v[i] = acc[i][j];
        if (apply_act_grad) {
#pragma unroll
          for (int i = 0; i < TM; ++i) v[i] *= act_grad_from_output(p[i], act);
        }
        store_rows<TM, F>(p, v);
      }
    }
  }
}

#define MOLANN_TM_DISPATCH(tm, CALL)                \
  switch (tm) {                                     \
    case 8: { constexpr int TM_ = 8; CALL; } break; \
    case 4: { constexpr int TM_ = 4; CALL; } break; \
    case 2: { constexpr int TM_ = 2; CALL; } break; \
    default: { constexpr int TM_ = 1; CALL; } break; \
  }

// Stage plan constants (indices, reference, feature program) into shared memory.
template <int NT>
__device__ __forceinline__ void stage_plan_consts(const DevPlan& p, const SmallLayout& lay, unsigned char* smem,
                                                  int tid) {
  int* aidx = reinterpret_cast<int*>(smem + lay.aidx_off);
  float* ref = reinterpret_cast<float*>(smem + lay.ref_off);
  int* ent = reinterpret_cast<int*>(smem + lay.ent_off);
  for (int i = tid; i < p.n_align; i += NT) aidx[i] = p.align_idx[i];
  for (int i = tid; i < 3 * p.n_align; i += NT) ref[i] = p.ref_x[i];
  for (int i = tid; i < ENTRY_INTS * p.n_entries; i += NT) ent[i] = p.entries[i];
}

struct TileOut {       // feature column writer: buffer [col][F], this thread's frame f
  float* base;
  int f;
  int F;
  __device__ __forceinline__ void operator()(int col, float v) { base[col * F + f] = v; }
};
struct TileGIn {       // feature cotangent reader
  const float* base;
  int f;
  int F;
  __device__ __forceinline__ float operator()(int col) const { return base[col * F + f]; }
  __device__ __forceinline__ void load2(int col, float& a, float& b) const { a = (*this)(col); b = (*this)(col + 1); }
  __device__ __forceinline__ void load3(int col, float& a, float& b, float& c) const {
    a = (*this)(col); b = (*this)(col + 1); c = (*this)(col + 2);
  }
};
struct RowAcc {        // plain accumulation into this thread's private gradient row
  float* row;
  __device__ __forceinline__ void operator()(int atom, V3 v) {
    float* q = row + 3 * atom;
    q[0] += v.x; q[1] += v.y; q[2] += v.z;
  }
};

// =============================================================================================
// Forward
// =============================================================================================
template <int F, int NT>
__global__ void __launch_bounds__(NT)
fused_small_forward_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ SmallLayout lay,
                           const float* __restrict__ x, float* __restrict__ y, long long L, int use_tma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  float* buf0 = reinterpret_cast<float*>(smem + lay.buf_off[0]);
  float* buf1 = reinterpret_cast<float*>(smem + lay.buf_off[1]);
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem + lay.mbar_off);

  stage_plan_consts<NT>(p, lay, smem, tid);
  for (int k = 0; k < p.n_layers; ++k) {      // transposed, zero-padded weights: Wt[i][o]
    const int K = p.dims[k], N = p.dims[k + 1], ld = lay.ldwt[k];
    float* Wt = reinterpret_cast<float*>(smem + lay.wt_off[k]);
    float* bs = reinterpret_cast<float*>(smem + lay.b_off[k]);
    const float* Wg = p.W[k];
    const float* bg = p.b[k];
    for (int idx = tid; idx < K * ld; idx += NT) {
      const int i = idx / ld, o = idx - i * ld;
      Wt[idx] = (o < N) ? Wg[(long long)o * K + i] : 0.f;
    }
    for (int o = tid; o < ld; o += NT) bs[o] = (o < N) ? bg[o] : 0.f;
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const long long ntiles = (L + F - 1) / F;
  const uint32_t tile_bytes = (uint32_t)F * (uint32_t)n3 * 4u;
  uint32_t phase = 0;
  auto is_tma_tile = [&](long long t) { return use_tma && (t + 1) * (long long)F <= L; };
  auto issue = [&](long long t) {
    if (tid == 0) {
      mbar_expect_tx(mbar, tile_bytes);
      bulk_g2s(xs, x + t * (long long)F * n3, tile_bytes, mbar);
    }
  };
  long long tile = blockIdx.x;
  if (tile < ntiles && is_tma_tile(tile)) issue(tile);

  for (; tile < ntiles; tile += gridDim.x) {
    const long long f_base = tile * (long long)F;
    const int nf = (int)((L - f_base) < (long long)F ? (L - f_base) : (long long)F);
    if (is_tma_tile(tile)) {
      mbar_wait(mbar, phase);
      phase ^= 1u;
    } else {
      const float* src = x + f_base * n3;
      for (int i = tid; i < nf * n3; i += NT) xs[i] = src[i];
      __syncthreads();
    }
    // ---- geometry: one thread per frame ----
    if (tid < F) {
      const int f = tid < nf ? tid : nf - 1;       // clamp: idle lanes recompute a valid frame
      const float* xf = xs + f * n3;
      Rigid rg;
      const bool aligned = p.n_align > 0;
      if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      TileOut out{buf0, tid, F};
      for (int e = 0; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    __syncthreads();
    const long long next = tile + gridDim.x;
    if (!lay.alias_xs && next < ntiles && is_tma_tile(next)) issue(next);   // overlaps the MLP below
    // ---- MLP ----
    const float* in = buf0;
    float* out = buf1;
    for (int k = 0; k < p.n_layers; ++k) {
      const bool last = (k == p.n_layers - 1);
      const float* Wt = reinterpret_cast<const float*>(smem + lay.wt_off[k]);
      const float* bs = reinterpret_cast<const float*>(smem + lay.b_off[k]);
      MOLANN_TM_DISPATCH(lay.tm_fwd[k],
                         (layer_forward<TM_, F, NT>(in, out, Wt, bs, lay.ldwt[k], p.dims[k], p.dims[k + 1], p.act,
                                                    last, y + f_base * p.dims[p.n_layers], nf, tid)));
      __syncthreads();
      const float* t = in;
      in = out;
      out = const_cast<float*>(t);
    }
    if (lay.alias_xs && next < ntiles && is_tma_tile(next)) issue(next);
  }
}

// =============================================================================================
// Backward (d<gy,y>/dx)
// =============================================================================================
// hidden activation h_k (k = 1 .. n_layers-1) lives in buffer slot(k); features / gfeat in slot 0.
__host__ __device__ inline int act_slot(int k) { return k == 1 ? 1 : (k == 2 ? 0 : k - 1); }

template <int F, int NT>
__global__ void __launch_bounds__(NT)
fused_small_backward_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ SmallLayout lay,
                            const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                            long long L, int use_tma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  float* gxs = reinterpret_cast<float*>(smem + lay.gxs_off);
  float* gys = reinterpret_cast<float*>(smem + lay.gys_off);
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem + lay.mbar_off);
  const int nl = p.n_layers;
  const int kout = p.dims[nl];

  stage_plan_consts<NT>(p, lay, smem, tid);
  for (int k = 0; k < nl; ++k) {
    const int K = p.dims[k], N = p.dims[k + 1];
    const float* Wg = p.W[k];
    // natural layout Wn[o][i], rows padded to a multiple of 8 outputs, ld = round_up(K, 8)
    {
      const int ld = lay.ldw[k], rows = round_up(N, 8);
      float* Wn = reinterpret_cast<float*>(smem + lay.w_off[k]);
      for (int idx = tid; idx < rows * ld; idx += NT) {
        const int o = idx / ld, i = idx - o * ld;
        Wn[idx] = (o < N && i < K) ? Wg[(long long)o * K + i] : 0.f;
      }
    }
    if (k < nl - 1) {      // transposed copy for the forward recompute (the last layer is not needed)
      const int ld = lay.ldwt[k];
      float* Wt = reinterpret_cast<float*>(smem + lay.wt_off[k]);
      float* bs = reinterpret_cast<float*>(smem + lay.b_off[k]);
      const float* bg = p.b[k];
      for (int idx = tid; idx < K * ld; idx += NT) {
        const int i = idx / ld, o = idx - i * ld;
        Wt[idx] = (o < N) ? Wg[(long long)o * K + i] : 0.f;
      }
      for (int o = tid; o < ld; o += NT) bs[o] = (o < N) ? bg[o] : 0.f;
    }
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const long long ntiles = (L + F - 1) / F;
  const uint32_t tile_bytes = (uint32_t)F * (uint32_t)n3 * 4u;
  uint32_t phase = 0;
  bool store_pending = false;
  auto is_tma_tile = [&](long long t) { return use_tma && (t + 1) * (long long)F <= L; };
  auto issue = [&](long long t) {
    if (tid == 0) {
      mbar_expect_tx(mbar, tile_bytes);
      bulk_g2s(xs, x + t * (long long)F * n3, tile_bytes, mbar);
    }
  };
  long long tile = blockIdx.x;
  if (tile < ntiles && is_tma_tile(tile)) issue(tile);

  for (; tile < ntiles; tile += gridDim.x) {
    const long long f_base = tile * (long long)F;
    const int nf = (int)((L - f_base) < (long long)F ? (L - f_base) : (long long)F);
    // cotangent tile gys[o][f]
    {
      const float* src = gy + f_base * kout;
      for (int i = tid; i < nf * kout; i += NT) {
        const int f = i / kout, o = i - f * kout;
        gys[o * F + f] = src[i];
      }
      for (int i = tid; i < (F - nf) * kout; i += NT) {      // zero cotangent for padded frames
        const int f = nf + i / kout, o = i % kout;
        gys[o * F + f] = 0.f;
      }
    }
    if (is_tma_tile(tile)) {
      mbar_wait(mbar, phase);
      phase ^= 1u;
    } else {
      const float* src = x + f_base * n3;
      for (int i = tid; i < nf * n3; i += NT) xs[i] = src[i];
    }
    __syncthreads();
    // ---- forward recompute: features + hidden activations ----
    Rigid rg;
    const bool aligned = p.n_align > 0;
    const int fclamp = tid < nf ? tid : nf - 1;
    const float* xf = xs + fclamp * n3;
    float* feat = reinterpret_cast<float*>(smem + lay.buf_off[0]);
    if (tid < F) {
      if (aligned) kabsch<1>(xf, aidx, ref, p.n_align, 0, rg);
      TileOut out{feat, tid, F};
      for (int e = 0; e < p.n_entries; ++e) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    __syncthreads();
    for (int k = 0; k < nl - 1; ++k) {
      const float* in = reinterpret_cast<const float*>(smem + lay.buf_off[k == 0 ? 0 : act_slot(k)]);
      float* out = reinterpret_cast<float*>(smem + lay.buf_off[act_slot(k + 1)]);
      const float* Wt = reinterpret_cast<const float*>(smem + lay.wt_off[k]);
      const float* bs = reinterpret_cast<const float*>(smem + lay.b_off[k]);
      MOLANN_TM_DISPATCH(lay.tm_fwd[k], (layer_forward<TM_, F, NT>(in, out, Wt, bs, lay.ldwt[k], p.dims[k],
                                                                   p.dims[k + 1], p.act, false, nullptr, nf, tid)));
      __syncthreads();
    }
    // ---- MLP backward, in place on the stored activations ----
    for (int k = nl - 1; k >= 0; --k) {
      const float* gz = (k == nl - 1) ? gys : reinterpret_cast<const float*>(smem + lay.buf_off[act_slot(k + 1)]);
      float* io = reinterpret_cast<float*>(smem + lay.buf_off[k == 0 ? 0 : act_slot(k)]);
      const float* Wn = reinterpret_cast<const float*>(smem + lay.w_off[k]);
      MOLANN_TM_DISPATCH(lay.tm_bwd[k], (layer_backward<TM_, F, NT>(gz, io, Wn, lay.ldw[k], p.dims[k + 1], p.dims[k],
                                                                    p.act, k > 0, tid)));
      __syncthreads();
    }
    // ---- feature + alignment backward into the gradient tile ----
    if (store_pending) {      // previous tile's bulk store must have finished reading gxs
      if (tid == 0) bulk_wait_read0();
      store_pending = false;
      __syncthreads();
    }
    for (int i = tid; i < F * n3; i += NT) gxs[i] = 0.f;
    __syncthreads();
    if (tid < F) {
      const float* gfeat = reinterpret_cast<const float*>(smem + lay.buf_off[0]);
      TileGIn gin{gfeat, tid, F};
      RowAcc acc{gxs + tid * n3};
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
    fence_proxy_async_smem();
    __syncthreads();
    // xs is free: prefetch the next tile while the gradient tile drains
    const long long next = tile + gridDim.x;
    if (next < ntiles && is_tma_tile(next)) issue(next);
    float* dst = gx + f_base * n3;
    if (is_tma_tile(tile)) {
      if (tid == 0) {
        bulk_s2g(dst, gxs, tile_bytes);
        bulk_commit();
      }
      store_pending = true;
    } else {
      for (int i = tid; i < nf * n3; i += NT) dst[i] = gxs[i];
    }
  }
  if (store_pending && tid == 0) bulk_wait0();
}

}  // namespace molann
