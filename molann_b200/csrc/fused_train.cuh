// fused_train.cuh -- one fused kernel for the autoencoder training step (BASELINE configs[3]).
//
// The reference ships no training loop; C4 is the composition its modules allow:
//   loss = mean((decoder(encoder(x)) - preprocessing(x))^2),  encoder = MolANN (molann/ann.py:567-624),
//   decoder = create_sequential_nn (ann.py:37-67), target = PreprocessingANN(x) (ann.py:553-565).
// Per tile of 128 frames a persistent CTA (one per SM) runs, without leaving the SM:
//   TMA bulk stage-in of x -> Kabsch + feature program (two lanes per frame) -> every Linear layer of encoder and
//   decoder forward (register-tiled FFMA, activations resident in shared memory as [width][frame] rows) -> residual,
//   loss and its cotangent -> for each layer, last to first: the weight / bias gradient of the layer (frames are the
//   contraction axis; a thread owns an 8 x 8 ... 4 x 4 block of dW, see the work shapes below) and the backward to the
//   layer input, in place.
// Parameter gradients accumulate in a per-CTA plane of the workspace (thread-private read-modify-write, L2 resident,
// no atomics: the result is deterministic); `train_reduce_kernel` sums the planes in a fixed order.  HBM traffic is the
// algorithmic 12 n bytes per frame; the step is bound by FP32 FMA work fed from shared memory (about 74 kFLOP per C4
// frame; DESIGN.md 3.8, profiles/r5_train_ab.txt).  Across ranks `train_allreduce_sgd_kernel` (end of this file) sums the
// flat vectors over NVLink peer memory and applies the SGD update in the same pass.
//
// Shared-memory rows have a stride of 132 floats: the dW contraction reads four rows per thread at the same frame
// quad, and with stride = 4 (mod 32) banks, rows r .. r+7 hit eight different bank quads.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "fused_small.cuh"
#include "tc.cuh"

namespace molann {

constexpr int TR_MAXL = 2 * MOLANN_MAX_LAYERS;   // encoder + decoder Linear layers
constexpr int TR_F = 128;                        // frames per tile
constexpr int TR_FS = TR_F + 4;                  // row stride in floats
constexpr int TR_NT = 256;                       // threads per CTA

struct TrainNet {                  // the chain encoder ++ decoder
  int nl, ne;                      // layers in total / in the encoder
  int act_enc, act_dec;
  int P;                           // trainable floats (flat gradient length; the loss sits at index P)
  int c[TR_MAXL + 1];              // widths: c[0] = d_feat, c[ne] = bottleneck, c[nl] = d_feat
  const float* W[TR_MAXL];
  const float* b[TR_MAXL];
  int gw[TR_MAXL], gb[TR_MAXL];    // offsets of dW_l / db_l in the flat gradient (torch parameter order)
};

struct TrainLayout {               // byte offsets into dynamic shared memory (host-computed)
  int a_off[TR_MAXL + 1];          // a_l = input of layer l ([round4(c[l])][TR_FS]); a_off[nl] = loss cotangent
  int w_off[TR_MAXL], b_off[TR_MAXL], ldk[TR_MAXL];
  int act_lo, act_bytes;           // the activation region (zeroed once: padding rows must stay 0)
  int xs_off;                      // coordinate tile, overlays the tail of the activation region
  int aidx_off, ref_off, ent_off, mbar_off, red_off, total_bytes;
  int fw[TR_MAXL], bw[TR_MAXL], dw[TR_MAXL];   // work shape of each phase of layer l (TR_WIDE ..., TR_DW44 ...)
};

__device__ __forceinline__ int tr_layer_act(const TrainNet& n, int l) {   // activation applied to layer l's output
  if (l == n.ne - 1 || l == n.nl - 1) return ACT_IDENTITY;              // create_sequential_nn: none after the last
  return l < n.ne ? n.act_enc : n.act_dec;
}

__device__ __forceinline__ float tr_act(float v, int act) {
  if (act == ACT_TANH) {                        // 1 - 2 / (1 + 2^(2 v log2 e)); saturates cleanly at +-1
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return fmaf(-2.0f, r, 1.0f);
  }
  if (act == ACT_SIGMOID) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
  }
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// ---- work shapes -------------------------------------------------------------------------------------------------
// Every phase between two CTA barriers should keep all eight warps busy, so the register tile of a work item shrinks
// with the layer: 4 frames x 8 columns while that gives 256 items (width > 56), 4 x 4 below, and for widths <= 4 (the
// bottleneck, a two-feature FeatureLayer) one thread pair per frame splitting the contraction instead of one warp
// doing all of it.  The host picks the shape per layer (TrainLayout::fw / bw / dw).
constexpr int TR_WIDE = 0, TR_HALF = 1, TR_NARROW = 2;              // fw / bw shapes
constexpr int TR_DW44 = 0, TR_DW42 = 1, TR_DW24 = 2, TR_DW22 = 3, TR_DWTHIN = 4;
constexpr int TR_DWB88 = 5, TR_DWB84 = 6, TR_DWB48 = 7, TR_DWB44 = 8;   // large tiles over frame quarters (tr_dw_big)

// acc[i][j] += sum_k A[k][f0 + i] * Wn[n0 + j][k]   (four frames x NO outputs, k in steps of four)
template <int NO>
__device__ __forceinline__ void tr_fwd_acc(const float* __restrict__ A, const float* __restrict__ Wn, int ldk, int K4,
                                           int f0, int n0, float (&acc)[4][NO]) {
  const float* ap = A + f0;
  const float* wp = Wn + n0 * ldk;
#pragma unroll 2
  for (int k = 0; k < K4; k += 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(ap);
    const float4 a1 = *reinterpret_cast<const float4*>(ap + TR_FS);
    const float4 a2 = *reinterpret_cast<const float4*>(ap + 2 * TR_FS);
    const float4 a3 = *reinterpret_cast<const float4*>(ap + 3 * TR_FS);
#pragma unroll
    for (int j = 0; j < NO; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(wp + j * ldk + k);
      acc[0][j] = fmaf(a3.x, w.w, fmaf(a2.x, w.z, fmaf(a1.x, w.y, fmaf(a0.x, w.x, acc[0][j]))));
      acc[1][j] = fmaf(a3.y, w.w, fmaf(a2.y, w.z, fmaf(a1.y, w.y, fmaf(a0.y, w.x, acc[1][j]))));
      acc[2][j] = fmaf(a3.z, w.w, fmaf(a2.z, w.z, fmaf(a1.z, w.y, fmaf(a0.z, w.x, acc[2][j]))));
      acc[3][j] = fmaf(a3.w, w.w, fmaf(a2.w, w.z, fmaf(a1.w, w.y, fmaf(a0.w, w.x, acc[3][j]))));
    }
    ap += 4 * TR_FS;
  }
}

// O[n][f] = act(sum_k A[k][f] Wn[n][k] + b[n]);  LAST: instead the residual against the features `feat`, the squared
// error of the valid frames (returned) and the loss cotangent two_scale * residual written to O
template <int NO, bool LAST>
__device__ __forceinline__ float tr_forward_layer(const float* __restrict__ A, float* __restrict__ O,
                                                  const float* __restrict__ Wn, const float* __restrict__ bs, int ldk,
                                                  int K4, int N, int act, const float* __restrict__ feat, int nf,
                                                  float two_scale, int tid) {
  const int nog = (N + NO - 1) / NO;
  float sq = 0.f;
  for (int item = tid; item < (TR_F / 4) * nog; item += TR_NT) {
    const int f0 = (item & 31) * 4, n0 = (item >> 5) * NO;
    float acc[4][NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) {
      const float bj = bs[n0 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][j] = bj;
    }
    tr_fwd_acc<NO>(A, Wn, ldk, K4, f0, n0, acc);
#pragma unroll
    for (int j = 0; j < NO; ++j) {
      if (n0 + j < N) {
        float4* dst = reinterpret_cast<float4*>(O + (n0 + j) * TR_FS + f0);
        if (LAST) {
          const float4 t = *reinterpret_cast<const float4*>(feat + (n0 + j) * TR_FS + f0);
          float e[4] = {acc[0][j] - t.x, acc[1][j] - t.y, acc[2][j] - t.z, acc[3][j] - t.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (f0 + i >= nf) e[i] = 0.f;
            sq = fmaf(e[i], e[i], sq);
            e[i] *= two_scale;
          }
          *dst = make_float4(e[0], e[1], e[2], e[3]);
        } else {
          *dst = make_float4(tr_act(acc[0][j], act), tr_act(acc[1][j], act), tr_act(acc[2][j], act),
                             tr_act(acc[3][j], act));
        }
      }
    }
  }
  return sq;
}

// the same layer for N <= 4: thread pair (f, kh) per frame, kh takes every other chunk of four k (rows 4 TR_FS apart
// are 16 banks apart: the two halves of a warp never collide), the pair is summed by one shuffle
template <bool LAST>
__device__ __forceinline__ float tr_forward_narrow(const float* __restrict__ A, float* __restrict__ O,
                                                   const float* __restrict__ Wn, const float* __restrict__ bs, int ldk,
                                                   int K4, int N, int act, const float* __restrict__ feat, int nf,
                                                   float two_scale, int tid) {
  const int f = tid >> 1, kh = tid & 1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 4 * kh; k < K4; k += 8) {
    const float a0 = A[k * TR_FS + f], a1 = A[(k + 1) * TR_FS + f], a2 = A[(k + 2) * TR_FS + f],
                a3 = A[(k + 3) * TR_FS + f];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(Wn + j * ldk + k);     // rows up to round8(N) exist (zeros)
      acc[j] = fmaf(a3, w.w, fmaf(a2, w.z, fmaf(a1, w.y, fmaf(a0, w.x, acc[j]))));
    }
  }
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float v = acc[j] + __shfl_xor_sync(0xffffffffu, acc[j], 1) + bs[j];
    if (j < N && (j & 1) == kh) {
      if (LAST) {
        float e = v - feat[j * TR_FS + f];
        if (f >= nf) e = 0.f;
        sq = fmaf(e, e, sq);
        O[j * TR_FS + f] = e * two_scale;
      } else {
        O[j * TR_FS + f] = tr_act(v, act);
      }
    }
  }
  return sq;
}

// IO[i][f] = (sum_o GZ[o][f] Wn[o][i]) * act'(IO[i][f])   -- in place on the layer input
template <int NI>
__device__ __forceinline__ void tr_backward_layer(const float* __restrict__ GZ, float* __restrict__ IO,
                                                  const float* __restrict__ Wn, int ldk, int N, int K, int act_prev,
                                                  int tid) {
  const int nig = (K + NI - 1) / NI;
  for (int item = tid; item < (TR_F / 4) * nig; item += TR_NT) {
    const int f0 = (item & 31) * 4, i0 = (item >> 5) * NI;
    float acc[4][NI];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j] = 0.f;
    const float* gp = GZ + f0;
    const float* wp = Wn + i0;
#pragma unroll 4
    for (int o = 0; o < N; ++o) {
      const float4 g = *reinterpret_cast<const float4*>(gp);
      float w[NI];
#pragma unroll
      for (int q = 0; q < NI / 4; ++q) {
        const float4 wq = *reinterpret_cast<const float4*>(wp + 4 * q);
        w[4 * q] = wq.x; w[4 * q + 1] = wq.y; w[4 * q + 2] = wq.z; w[4 * q + 3] = wq.w;
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        acc[0][j] = fmaf(g.x, w[j], acc[0][j]);
        acc[1][j] = fmaf(g.y, w[j], acc[1][j]);
        acc[2][j] = fmaf(g.z, w[j], acc[2][j]);
        acc[3][j] = fmaf(g.w, w[j], acc[3][j]);
      }
      gp += TR_FS;
      wp += ldk;
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      if (i0 + j < K) {
        float4* p = reinterpret_cast<float4*>(IO + (i0 + j) * TR_FS + f0);
        const float4 h = *p;
        *p = make_float4(acc[0][j] * act_grad_from_output(h.x, act_prev), acc[1][j] * act_grad_from_output(h.y, act_prev),
                         acc[2][j] * act_grad_from_output(h.z, act_prev), acc[3][j] * act_grad_from_output(h.w, act_prev));
      }
    }
  }
}

// the same for K <= 4 (the cotangent of the bottleneck): thread pair (f, oh), oh takes every other chunk of four o
__device__ __forceinline__ void tr_backward_narrow(const float* __restrict__ GZ, float* __restrict__ IO,
                                                   const float* __restrict__ Wn, int ldk, int N, int K, int act_prev,
                                                   int tid) {
  const int f = tid >> 1, oh = tid & 1;
  const int N4 = round_up(N, 4);                      // GZ rows and W rows up to N4 exist and are zero
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int o = 4 * oh; o < N4; o += 8) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float g = GZ[(o + r) * TR_FS + f];
      const float4 w = *reinterpret_cast<const float4*>(Wn + (o + r) * ldk);
      acc[0] = fmaf(g, w.x, acc[0]);
      acc[1] = fmaf(g, w.y, acc[1]);
      acc[2] = fmaf(g, w.z, acc[2]);
      acc[3] = fmaf(g, w.w, acc[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float v = acc[j] + __shfl_xor_sync(0xffffffffu, acc[j], 1);
    if (j < K && (j & 1) == oh) {
      float* p = IO + j * TR_FS + f;
      *p = v * act_grad_from_output(*p, act_prev);
    }
  }
}

// dW[o][i] += sum_f GZ[o][f] A[i][f],  db[o] += sum_f GZ[o][f].  A thread owns rows o = to + TO r (r < RO) and
// i = ti + TI r (r < RI), interleaved so that the eight `to` of a warp read eight different bank quads; a warp covers
// 8 x 4 threads.  The contraction runs on the packed fp32 pipe: the (even frame, odd frame) partial sums of one
// dW[o][i] share a register pair, both operands of an FFMA2 are the natural pairs of the float4 rows.
// The running sums live in this CTA's plane of the workspace; nobody else touches it.
template <int RO, int RI>
__device__ __forceinline__ void tr_dw_layer(const float* __restrict__ GZ, const float* __restrict__ A, int N, int K,
                                            float* __restrict__ pw, float* __restrict__ pb, bool first, int tid) {
  const int TO = (N + RO - 1) / RO, TI = (K + RI - 1) / RI;
  const int nbo = (TO + 7) >> 3, nbi = (TI + 3) >> 2;
  const int warp = tid >> 5, lane = tid & 31;
  for (int blk = warp; blk < nbo * nbi; blk += TR_NT / 32) {
    const int to = (blk % nbo) * 8 + (lane & 7), ti = (blk / nbo) * 4 + (lane >> 3);
    if (to >= TO || ti >= TI) continue;
    const bool with_bias = (ti == 0);
    // the running sums are fetched now and only needed after the frame loop: their L2 latency hides behind it
    unsigned long long acc[RO][RI], sb[RO];
    float oldw[RO][RI], oldb[RO];
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {
      const int o = to + TO * ro;
      oldb[ro] = (with_bias && !first && o < N) ? __ldcg(pb + o) : 0.f;
      sb[ro] = f2_pack(0.f, 0.f);
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) {
        const int i = ti + TI * ri;
        oldw[ro][ri] = (!first && o < N && i < K) ? __ldcg(pw + o * K + i) : 0.f;
        acc[ro][ri] = f2_pack(0.f, 0.f);
      }
    }
    // rows past the layer's width are read as whatever follows them; clamp to a valid row (their sums are dropped)
    const float* gp[RO];
    const float* ap[RI];
#pragma unroll
    for (int r = 0; r < RO; ++r) gp[r] = GZ + (to + TO * r < N ? to + TO * r : to) * TR_FS;
#pragma unroll
    for (int r = 0; r < RI; ++r) ap[r] = A + (ti + TI * r < K ? ti + TI * r : ti) * TR_FS;
#pragma unroll 2
    for (int fq = 0; fq < TR_F; fq += 4) {
      ulonglong2 g[RO], a[RI];
#pragma unroll
      for (int r = 0; r < RO; ++r) g[r] = *reinterpret_cast<const ulonglong2*>(gp[r] + fq);
#pragma unroll
      for (int r = 0; r < RI; ++r) a[r] = *reinterpret_cast<const ulonglong2*>(ap[r] + fq);
#pragma unroll
      for (int ro = 0; ro < RO; ++ro)
#pragma unroll
        for (int ri = 0; ri < RI; ++ri)
          acc[ro][ri] = f2_fma(g[ro].y, a[ri].y, f2_fma(g[ro].x, a[ri].x, acc[ro][ri]));
      if (with_bias) {
#pragma unroll
        for (int ro = 0; ro < RO; ++ro) sb[ro] = f2_add(f2_add(sb[ro], g[ro].x), g[ro].y);
      }
    }
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {
      const int o = to + TO * ro;
      if (o < N) {
        float lo, hi;
        if (with_bias) {
          f2_unpack(sb[ro], lo, hi);
          __stcg(pb + o, oldb[ro] + (lo + hi));
        }
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) {
          const int i = ti + TI * ri;
          f2_unpack(acc[ro][ri], lo, hi);
          if (i < K) __stcg(pw + o * K + i, oldw[ro][ri] + (lo + hi));
        }
      }
    }
  }
}

// The large-tile form of the same phase (layers at least ~ 28 wide on both sides, K <= 8 RI).  The contractions are
// bound by the bytes shared memory returns to the register file (profiles/r5_train_ab.txt), i.e. by FMAs per loaded
// float: an RO x RI = 8 x 8 tile needs half the loads of a 4 x 4 one.  To keep every lane busy with tiles that large the
// four lane octets of a warp take the four frame quarters of the tile: lane = (t, q), columns i = t + TI r (r < RI)
// over frames [32 q, 32 q + 32); warp w owns the row groups wo = w, w + 8, ... (o = wo + TO r, r < RO; its GZ rows
// are warp-wide broadcasts).  The four partial sums of a dW entry meet by recursive halving over the two lane bits of q
// (RO RI / 2 + RO RI / 4 shuffles instead of 2 RO RI), after which every lane holds RO RI / 4 finished entries; the
// eight t-lanes of an octet hold eight consecutive columns of a dW row, one 32-byte sector of the plane.
template <int RO, int RI>
__device__ __forceinline__ void tr_dw_big(const float* __restrict__ GZ, const float* __restrict__ A, int N, int K,
                                          float* __restrict__ pw, float* __restrict__ pb, bool first, int tid) {
  constexpr int V = RO * RI;
  const int TO = (N + RO - 1) / RO, TI = (K + RI - 1) / RI;      // TI <= 8 (host)
  const int warp = tid >> 5, lane = tid & 31, t = lane & 7, q = lane >> 3;
  const bool col_ok = t < TI;
  const bool hi2 = (q & 2) != 0, hi1 = (q & 1) != 0;
  const int base = (hi2 ? V / 2 : 0) + (hi1 ? V / 4 : 0);       // first of the V / 4 entries this lane finishes
  const float* ap[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) ap[r] = A + ((col_ok && t + TI * r < K) ? t + TI * r : 0) * TR_FS + 32 * q;
  for (int wo = warp; wo < TO; wo += TR_NT / 32) {
    float acc[V], sb[RO];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
#pragma unroll
    for (int r = 0; r < RO; ++r) sb[r] = 0.f;
    const float* gp[RO];
#pragma unroll
    for (int r = 0; r < RO; ++r) gp[r] = GZ + (wo + TO * r < N ? wo + TO * r : wo) * TR_FS + 32 * q;
#pragma unroll 1
    for (int fq = 0; fq < 32; fq += 4) {
      float4 g[RO], a[RI];
#pragma unroll
      for (int r = 0; r < RO; ++r) g[r] = *reinterpret_cast<const float4*>(gp[r] + fq);
#pragma unroll
      for (int r = 0; r < RI; ++r) a[r] = *reinterpret_cast<const float4*>(ap[r] + fq);
#pragma unroll
      for (int ro = 0; ro < RO; ++ro) {
#pragma unroll
        for (int ri = 0; ri < RI; ++ri)
          acc[ro * RI + ri] = fmaf(g[ro].w, a[ri].w, fmaf(g[ro].z, a[ri].z, fmaf(g[ro].y, a[ri].y,
                                   fmaf(g[ro].x, a[ri].x, acc[ro * RI + ri]))));
        sb[ro] += (g[ro].x + g[ro].y) + (g[ro].z + g[ro].w);
      }
    }
    // the running sums of this lane's final entries: issued now, needed after the shuffles (the cache-hinted accesses
    // are volatile asm -- interleaved with the stores each would wait a full L2 round trip)
    float oldv[V / 4], oldb[RO];
#pragma unroll
    for (int v = 0; v < V / 4; ++v) {
      const int idx = base + v, o = wo + TO * (idx / RI), i = t + TI * (idx % RI);
      oldv[v] = (!first && col_ok && o < N && i < K) ? __ldcg(pw + o * K + i) : 0.f;
    }
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {
      const int o = wo + TO * ro;
      oldb[ro] = (!first && lane == 0 && o < N) ? __ldcg(pb + o) : 0.f;
    }
    // frame quarters meet: xor 16 halves the entries a lane is responsible for, xor 8 halves them again
    float h1[V / 2], h2[V / 4];
#pragma unroll
    for (int v = 0; v < V / 2; ++v) {
      const float mine = hi2 ? acc[v + V / 2] : acc[v], other = hi2 ? acc[v] : acc[v + V / 2];
      h1[v] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
    }
#pragma unroll
    for (int v = 0; v < V / 4; ++v) {
      const float mine = hi1 ? h1[v + V / 4] : h1[v], other = hi1 ? h1[v] : h1[v + V / 4];
      h2[v] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
    }
#pragma unroll
    for (int v = 0; v < V / 4; ++v) {
      const int idx = base + v, o = wo + TO * (idx / RI), i = t + TI * (idx % RI);
      if (col_ok && o < N && i < K) __stcg(pw + o * K + i, oldv[v] + h2[v]);
    }
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {                  // every lane of the warp holds the same GZ rows: lane 0 writes db
      float b = sb[ro] + __shfl_xor_sync(0xffffffffu, sb[ro], 16);
      b += __shfl_xor_sync(0xffffffffu, b, 8);
      const int o = wo + TO * ro;
      if (lane == 0 && o < N) __stcg(pb + o, oldb[ro] + b);
    }
  }
}

// the same for min(N, K) <= 4: one thread per row w of the wide side contracts it with the (at most four) rows of the
// thin side over all frames; the thin rows are warp-wide broadcasts
__device__ __forceinline__ void tr_dw_thin(const float* __restrict__ GZ, const float* __restrict__ A, int N, int K,
                                           float* __restrict__ pw, float* __restrict__ pb, bool first, int tid) {
  const bool n_thin = N <= K;                          // thin rows come from GZ (outputs), wide rows from A -- or not
  const float* thin = n_thin ? GZ : A;
  const float* wide = n_thin ? A : GZ;
  const int T = n_thin ? N : K, Wd = n_thin ? K : N;
  for (int w = tid; w < Wd; w += TR_NT) {
    unsigned long long acc[4], sb = f2_pack(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = f2_pack(0.f, 0.f);
    const float* wp = wide + w * TR_FS;
#pragma unroll 2
    for (int fq = 0; fq < TR_F; fq += 4) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(wp + fq);
#pragma unroll
      for (int r = 0; r < 4; ++r) {                    // rows up to round4(T) exist (zeros)
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(thin + r * TR_FS + fq);
        acc[r] = f2_fma(v.y, t.y, f2_fma(v.x, t.x, acc[r]));
      }
      sb = f2_add(f2_add(sb, v.x), v.y);
    }
    float lo, hi;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < T) {
        float* q = n_thin ? pw + r * K + w : pw + w * K + r;
        f2_unpack(acc[r], lo, hi);
        __stcg(q, (first ? 0.f : __ldcg(q)) + (lo + hi));
      }
    }
    if (!n_thin) {                                     // wide rows are the outputs: their frame sums are db
      f2_unpack(sb, lo, hi);
      __stcg(pb + w, (first ? 0.f : __ldcg(pb + w)) + (lo + hi));
    }
  }
  if (n_thin && tid >= TR_NT - 32) {                   // db of the thin outputs: the last warp, one row at a time
    const int lane = tid & 31;
    for (int r = 0; r < T; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(thin + r * TR_FS + 4 * lane);
      float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      if (lane == 0) __stcg(pb + r, (first ? 0.f : __ldcg(pb + r)) + s);
    }
  }
}

__global__ void __launch_bounds__(TR_NT, 1)
fused_train_kernel(const __grid_constant__ DevPlan p, const __grid_constant__ TrainNet net,
                   const __grid_constant__ TrainLayout lay, const float* __restrict__ x, long long L, float loss_scale,
                   float* __restrict__ planes, int use_tma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x;
  const int n3 = 3 * p.n_inp;
  float* xs = reinterpret_cast<float*>(smem + lay.xs_off);
  const int* aidx = reinterpret_cast<const int*>(smem + lay.aidx_off);
  const float* ref = reinterpret_cast<const float*>(smem + lay.ref_off);
  const int* ent = reinterpret_cast<const int*>(smem + lay.ent_off);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem + lay.mbar_off);
  float* red = reinterpret_cast<float*>(smem + lay.red_off);
  float* plane = planes + (size_t)blockIdx.x * (size_t)(net.P + 1);

  // ---- once per CTA: constants, zero-padded natural weights Wn[round8(N)][ldk], zeroed activation rows ----
  {
    int* aidx_w = reinterpret_cast<int*>(smem + lay.aidx_off);
    float* ref_w = reinterpret_cast<float*>(smem + lay.ref_off);
    int* ent_w = reinterpret_cast<int*>(smem + lay.ent_off);
    for (int i = tid; i < p.n_align; i += TR_NT) aidx_w[i] = p.align_idx[i];
    for (int i = tid; i < 3 * p.n_align; i += TR_NT) ref_w[i] = p.ref_x[i];
    for (int i = tid; i < ENTRY_INTS * p.n_entries; i += TR_NT) ent_w[i] = p.entries[i];
  }
  for (int l = 0; l < net.nl; ++l) {
    const int K = net.c[l], N = net.c[l + 1], ld = lay.ldk[l], N8 = round_up(N, 8);
    float* Wn = reinterpret_cast<float*>(smem + lay.w_off[l]);
    float* bs = reinterpret_cast<float*>(smem + lay.b_off[l]);
    const float* Wg = net.W[l];
    const float* bg = net.b[l];
    for (int idx = tid; idx < N8 * ld; idx += TR_NT) {
      const int o = idx / ld, i = idx - o * ld;
      Wn[idx] = (o < N && i < K) ? Wg[(long long)o * K + i] : 0.f;
    }
    for (int o = tid; o < N8; o += TR_NT) bs[o] = (o < N) ? bg[o] : 0.f;
  }
  {
    float4* z = reinterpret_cast<float4*>(smem + lay.act_lo);
    for (int i = tid; i < lay.act_bytes / 16; i += TR_NT) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const long long ntiles = (L + TR_F - 1) / TR_F;
  const uint32_t tile_bytes = (uint32_t)TR_F * (uint32_t)n3 * 4u;
  const float two_scale = 2.0f * loss_scale;
  uint32_t phase = 0;
  float sq = 0.f;
  float* feat = reinterpret_cast<float*>(smem + lay.a_off[0]);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const bool first = (tile == (long long)blockIdx.x);
    const long long f_base = tile * (long long)TR_F;
    const int nf = (int)((L - f_base) < (long long)TR_F ? (L - f_base) : (long long)TR_F);
    // ---- stage the coordinate tile (it overlays activation rows that are dead at this point) ----
    if (use_tma && nf == TR_F) {
      if (tid == 0) {
        fence_proxy_async_smem();
        mbar_expect_tx(mbar, tile_bytes);
        bulk_g2s(xs, x + f_base * n3, tile_bytes, mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1u;
    } else {
      const float* src = x + f_base * n3;
      for (int i = tid; i < nf * n3; i += TR_NT) xs[i] = src[i];
      __syncthreads();
    }
    // ---- geometry: two lanes per frame ----
    {
      const int f = tid >> 1, sub = tid & 1;
      const float* xf = xs + (f < nf ? f : nf - 1) * n3;     // idle frame slots recompute a valid frame (masked below)
      Rigid rg;
      const bool aligned = p.n_align > 0;
      if (aligned) kabsch<2>(xf, aidx, ref, p.n_align, sub, rg);
      TileOut out{feat, f, TR_FS};
      for (int e = sub; e < p.n_entries; e += 2) {
        const Entry en = load_entry(ent + ENTRY_INTS * e);
        feature_forward(en, xf, aligned, rg, p.use_angle, out);
      }
    }
    __syncthreads();
    // the overlay is dead from here on; its rows must read as zero padding again where they are padding
    if (lay.xs_off < lay.act_lo + lay.act_bytes) {
      float4* z = reinterpret_cast<float4*>(xs);
      for (int i = tid; i < (int)(tile_bytes / 16); i += TR_NT) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();
    }
    // ---- forward through encoder and decoder ----
    for (int l = 0; l < net.nl; ++l) {
      const float* A = reinterpret_cast<const float*>(smem + lay.a_off[l]);
      float* O = reinterpret_cast<float*>(smem + lay.a_off[l + 1]);
      const float* Wn = reinterpret_cast<const float*>(smem + lay.w_off[l]);
      const float* bs = reinterpret_cast<const float*>(smem + lay.b_off[l]);
      const int K4 = round_up(net.c[l], 4), N = net.c[l + 1], ld = lay.ldk[l], act = tr_layer_act(net, l);
      const int shape = lay.fw[l];
      if (l < net.nl - 1) {
        if (shape == TR_WIDE) tr_forward_layer<8, false>(A, O, Wn, bs, ld, K4, N, act, feat, nf, two_scale, tid);
        else if (shape == TR_HALF) tr_forward_layer<4, false>(A, O, Wn, bs, ld, K4, N, act, feat, nf, two_scale, tid);
        else tr_forward_narrow<false>(A, O, Wn, bs, ld, K4, N, act, feat, nf, two_scale, tid);
      } else {
        if (shape == TR_WIDE) sq += tr_forward_layer<8, true>(A, O, Wn, bs, ld, K4, N, act, feat, nf, two_scale, tid);
        else if (shape == TR_HALF) sq += tr_forward_layer<4, true>(A, O, Wn, bs, ld, K4, N, act, feat, nf, two_scale, tid);
        else sq += tr_forward_narrow<true>(A, O, Wn, bs, ld, K4, N, act, feat, nf, two_scale, tid);
      }
      __syncthreads();
    }
    // ---- backward: parameter gradients of layer l, then the cotangent of its input (in place) ----
    for (int l = net.nl - 1; l >= 0; --l) {
      const float* GZ = reinterpret_cast<const float*>(smem + lay.a_off[l + 1]);
      float* A = reinterpret_cast<float*>(smem + lay.a_off[l]);
      const int K = net.c[l], N = net.c[l + 1];
      float* pw = plane + net.gw[l];
      float* pb = plane + net.gb[l];
      switch (lay.dw[l]) {
        case TR_DW44: tr_dw_layer<4, 4>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DW42: tr_dw_layer<4, 2>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DW24: tr_dw_layer<2, 4>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DW22: tr_dw_layer<2, 2>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DWB88: tr_dw_big<8, 8>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DWB84: tr_dw_big<8, 4>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DWB48: tr_dw_big<4, 8>(GZ, A, N, K, pw, pb, first, tid); break;
        case TR_DWB44: tr_dw_big<4, 4>(GZ, A, N, K, pw, pb, first, tid); break;
        default: tr_dw_thin(GZ, A, N, K, pw, pb, first, tid); break;
      }
      if (l > 0) {
        __syncthreads();
        const float* Wn = reinterpret_cast<const float*>(smem + lay.w_off[l]);
        const int act_prev = tr_layer_act(net, l - 1);
        if (lay.bw[l] == TR_WIDE) tr_backward_layer<8>(GZ, A, Wn, lay.ldk[l], N, K, act_prev, tid);
        else if (lay.bw[l] == TR_HALF) tr_backward_layer<4>(GZ, A, Wn, lay.ldk[l], N, K, act_prev, tid);
        else tr_backward_narrow(GZ, A, Wn, lay.ldk[l], N, K, act_prev, tid);
      }
      __syncthreads();
    }
  }
  // ---- this CTA's share of the loss ----
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < TR_NT / 32; ++w) t += red[w];
    plane[net.P] = t * loss_scale;
  }
}

// flat[p] = sum over the CTA planes in a fixed order (deterministic): block (64, 8), eight interleaved groups of planes per
// entry, then the eight partial sums pairwise
constexpr int TR_RED_X = 64, TR_RED_Y = 8;
__global__ void __launch_bounds__(TR_RED_X * TR_RED_Y)
train_reduce_kernel(const float* __restrict__ planes, int n_planes, int n, float* __restrict__ flat) {
  __shared__ float part[TR_RED_Y][TR_RED_X];
  const int i = blockIdx.x * TR_RED_X + threadIdx.x, g = threadIdx.y;
  float s = 0.f;
  if (i < n) {
#pragma unroll 4
    for (int c = g; c < n_planes; c += TR_RED_Y) s += __ldcg(planes + (size_t)c * n + i);
  }
  part[g][threadIdx.x] = s;
  __syncthreads();
  if (g == 0 && i < n) {
    const int x = threadIdx.x;
    flat[i] = ((part[0][x] + part[1][x]) + (part[2][x] + part[3][x])) + ((part[4][x] + part[5][x]) + (part[6][x] + part[7][x]));
  }
}

// plain SGD over up to 2 TR_MAXL parameter tensors laid out like the flat gradient: p -= lr * g
struct SgdTable {
  float* ptr[2 * TR_MAXL];
  int end[2 * TR_MAXL];            // exclusive prefix ends in the flat gradient
  int n;
};
__global__ void train_sgd_kernel(const __grid_constant__ SgdTable tab, const float* __restrict__ flat, float lr,
                                 int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int t = 0;
  while (t < tab.n - 1 && i >= tab.end[t]) ++t;
  const int start = t ? tab.end[t - 1] : 0;
  float* q = tab.ptr[t] + (i - start);
  *q = fmaf(-lr, flat[i], *q);
}

// ---- data-parallel step: one-shot allreduce over NVLink peer memory fused with the SGD update -----------------------------
// Every rank holds the flat vector of its shard (gradients + loss).  Instead of an NCCL allreduce followed by the SGD launch,
// ONE kernel per rank: (1) publish the local vector in this rank's symmetric buffer (slot = step parity), (2) the last CTA to
// finish signals every peer (release store of the step number into the peer's flag row, over NVLink), (3) every CTA waits
// for all peers' flags, reads the W published vectors through peer pointers, sums them in RANK ORDER -- every rank computes
// bit-identical sums, so the replicas never drift -- writes the global vector and applies p -= lr * g.
// The step number lives in device memory and is advanced by the kernel itself, so a CUDA graph can replay the launch.
// A rank cannot run two steps ahead of a peer (its step s + 1 needs the peer's step-s + 1 flag, which the peer raises after
// it has finished reading step s), so two slots are enough.
constexpr int TR_MAX_PEERS = 8;
struct PeerTable {
  float* buf[TR_MAX_PEERS];            // peer r's symmetric buffer: 2 slots of n floats, then the flag row
  unsigned* flag[TR_MAX_PEERS];        // peer r's flag row: flag[r][q] = last step rank q has published
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {          // never from a stale cache line
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// state[0] = number of the step this launch performs (starts at 1), state[1] / state[2] = CTA arrival counters.
// The grid must be co-resident (host: at most one CTA per SM).
__global__ void __launch_bounds__(256)
train_allreduce_sgd_kernel(const __grid_constant__ PeerTable peers, const __grid_constant__ SgdTable tab,
                           const float* __restrict__ flat_local, float* __restrict__ flat_global, int n, int total,
                           float lr, unsigned* __restrict__ state) {
  const unsigned step = *reinterpret_cast<volatile unsigned*>(state);
  const int slot = (int)(step & 1u) * n;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  float* own = peers.buf[peers.rank] + slot;
  for (int i = gtid; i < n; i += gstride) own[i] = flat_local[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(state + 1, 1u) == gridDim.x - 1) {               // every CTA of this rank has published
      __threadfence_system();
      for (int r = 0; r < peers.world; ++r) st_release_sys(peers.flag[r] + peers.rank, step);
    }
    for (int r = 0; r < peers.world; ++r)
      while (ld_acquire_sys(peers.flag[peers.rank] + r) < step) __nanosleep(64);
  }
  __syncthreads();
  for (int i = gtid; i < n; i += gstride) {
    float sum = 0.f;
    for (int r = 0; r < peers.world; ++r) sum += ld_peer(peers.buf[r] + slot + i);
    flat_global[i] = sum;
    if (i < total && lr != 0.f) {
      int t = 0;
      while (t < tab.n - 1 && i >= tab.end[t]) ++t;
      float* q = tab.ptr[t] + (i - (t ? tab.end[t - 1] : 0));
      *q = fmaf(-lr, sum, *q);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 2, 1u) == gridDim.x - 1) {               // last CTA out: re-arm for the next launch
      state[1] = 0u;
      state[2] = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(state) = step + 1u;
    }
  }
}

}  // namespace molann
