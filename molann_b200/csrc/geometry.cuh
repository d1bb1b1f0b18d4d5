// geometry.cuh -- per-frame geometry of the molann hot path as device functions.
//
// One frame is handled by a *group* of G lanes (G = 1: thread-per-frame on an smem-staged tile,
// G = 32: warp-per-frame straight from global memory).  All loops are lane-strided and all
// reductions are xor-butterflies of width G, so the same code serves both mappings.
//
// Math (reference molann/ann.py):
//   alignment  :179-197  c = mean(x[A]);  H = sum_k (x[A_k]-c)^T Y_k;  R = argmax_R tr(R^T H), det R = +1
//              solved as the dominant eigenvector of Horn's 4x4 quaternion matrix (cyclic Jacobi in
//              registers) -- identical to U diag(1,1,sign det(U V^T)) V^T of the reference's SVD route.
//   features   :323-354  angle / bond / dihedral / position.  Internal coordinates are rigid-motion
//              invariant, so they are evaluated on the RAW coordinates; only position features go
//              through (c, R).
//   backward   autograd of the above in closed form (no SVD derivative):  with G_j = dL/dz_j,
//              M = sum_j (x_j-c)^T G_j, P = R^T H, solve (tr(P) I - P) w = vee(R^T M - M^T R),
//              dL/dH = R [w]x,  gx_j = G_j R^T,  gx_{A_k} += Y_k (dL/dH)^T - (1/n_a) sum_j G_j R^T.
#pragma once
#include <stdint.h>
#ifdef MOLANN_HOST_EMULATION
// tests/host_emulation.cpp compiles these device functions for the CPU (G = 1 only) so the closed-form
// math can be checked against the oracle without a GPU.  Never used by the product.
#include <cmath>
#define __device__
#define __forceinline__ inline
#define __restrict__
inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
inline bool __all_sync(unsigned, bool p) { return p; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __ldg(const float* p) { return *p; }
inline float __fdividef(float a, float b) { return a / b; }
#else
#include <cuda_runtime.h>
#endif

namespace molann {

enum { FEAT_ANGLE = 0, FEAT_BOND = 1, FEAT_DIHEDRAL = 2, FEAT_POSITION = 3 };
enum { ACT_TANH = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_IDENTITY = 3 };
constexpr int ENTRY_INTS = 6;

template <int G>
__device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// 4x4 symmetric eigenproblem: cyclic Jacobi, everything in registers (indices are compile-time).
// ------------------------------------------------------------------------------------------------
// Jacobi rotation in the (P,Q) plane: (c, s) with c^2 + s^2 = 1 to fp32 rounding, chosen to zero a[P][Q].
//   t = sgn(d) 2 a_pq / (|d| + sqrt(d^2 + 4 a_pq^2)),  d = a_qq - a_pp   (the smaller root of t^2 + (d/a_pq) t - 1)
// Approximate MUFU ops (rsqrt / rcp) only perturb the ANGLE (convergence stays quadratic); orthogonality of
// the rotation rests on c = rsqrt(1 + t^2), which gets one Newton step so that c^2 (1 + t^2) = 1 to rounding.
__device__ __forceinline__ void jacobi_params(float app, float aqq, float apq, float& c, float& s, float& t) {
  const float d = aqq - app;
  const float r2 = fmaf(d, d, 4.0f * apq * apq);
  float t_ = 0.0f;
  if (r2 > 0.0f) {
    const float r = r2 * rsqrtf(r2);
    t_ = __fdividef(copysignf(2.0f * apq, d * apq), fabsf(d) + r);     // sign(t) = sign(d * a_pq)
    if (d == 0.0f) t_ = copysignf(1.0f, apq);
  }
  const float w = fmaf(t_, t_, 1.0f);
  float c_ = rsqrtf(w);
  c_ = c_ * fmaf(-0.5f * w, c_ * c_, 1.5f);
  t = t_;
  c = c_;
  s = t_ * c_;
}

// apply the rotation of plane (P,Q) to the symmetric matrix (upper triangle in `a`) and accumulate it in v
template <int P, int Q, int R1, int R2>
__device__ __forceinline__ void jacobi_apply(float (&a)[4][4], float (&v)[4][4], float c, float s, float t) {
  const float apq = a[P][Q];
  a[P][P] = fmaf(-t, apq, a[P][P]);
  a[Q][Q] = fmaf(t, apq, a[Q][Q]);
  a[P][Q] = 0.0f;
  {
    float& x1 = (R1 < P) ? a[R1][P] : a[P][R1];
    float& y1 = (R1 < Q) ? a[R1][Q] : a[Q][R1];
    const float g = x1, h = y1;
    x1 = fmaf(c, g, -s * h);
    y1 = fmaf(s, g, c * h);
    float& x2 = (R2 < P) ? a[R2][P] : a[P][R2];
    float& y2 = (R2 < Q) ? a[R2][Q] : a[Q][R2];
    const float g2 = x2, h2 = y2;
    x2 = fmaf(c, g2, -s * h2);
    y2 = fmaf(s, g2, c * h2);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float g = v[r][P], h = v[r][Q];
    v[r][P] = fmaf(c, g, -s * h);
    v[r][Q] = fmaf(s, g, c * h);
  }
}

// Two rotations on DISJOINT planes (P,Q) and (R,S): their parameters depend on disjoint matrix entries, so
// both long-latency parameter chains are issued back to back (ILP 2) before the updates are applied.
template <int P, int Q, int R, int S>
__device__ __forceinline__ void jacobi_pair(float (&a)[4][4], float (&v)[4][4]) {
  float c1, s1, t1, c2, s2, t2;
  jacobi_params(a[P][P], a[Q][Q], a[P][Q], c1, s1, t1);
  jacobi_params(a[R][R], a[S][S], a[R][S], c2, s2, t2);
  jacobi_apply<P, Q, R, S>(a, v, c1, s1, t1);
  jacobi_apply<R, S, P, Q>(a, v, c2, s2, t2);
}

// Dominant (largest-eigenvalue) unit eigenvector of the symmetric matrix whose upper triangle is `a`.
// Parallel-ordered cyclic Jacobi: a sweep is {(0,1),(2,3)}, {(0,2),(1,3)}, {(0,3),(1,2)}.
// Must be called by all 32 lanes of a warp (uses a warp vote for a uniform exit).
__device__ __forceinline__ void dominant_eigvec4(float (&a)[4][4], float (&q)[4], bool active = true) {
  float v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = (i == j) ? 1.0f : 0.0f;
  // Each lane latches its own convergence, so a frame's result never depends on its warp-mates (frames are
  // bitwise independent of batch composition / sharding); the vote only decides when the warp leaves the loop.
  bool done = !active;          // lanes that do not need the result never hold the warp in the loop
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    if (!done) {
      const float off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[0][3] * a[0][3] + a[1][2] * a[1][2] +
                        a[1][3] * a[1][3] + a[2][3] * a[2][3];
      const float dg = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2] + a[3][3] * a[3][3];
      done = !(off > 4e-14f * dg);
    }
    if (__all_sync(0xffffffffu, done)) break;
    if (!done) {
      jacobi_pair<0, 1, 2, 3>(a, v);
      jacobi_pair<0, 2, 1, 3>(a, v);
      jacobi_pair<0, 3, 1, 2>(a, v);
    }
  }
  int best = 0;
  float lam = a[0][0];
  if (a[1][1] > lam) { lam = a[1][1]; best = 1; }
  if (a[2][2] > lam) { lam = a[2][2]; best = 2; }
  if (a[3][3] > lam) { lam = a[3][3]; best = 3; }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    q[r] = (best == 0) ? v[r][0] : (best == 1) ? v[r][1] : (best == 2) ? v[r][2] : v[r][3];
  const float inv = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
#pragma unroll
  for (int r = 0; r < 4; ++r) q[r] *= inv;
}


// ------------------------------------------------------------------------------------------------
// Fast path for the same eigenvector: characteristic-polynomial Newton (Theobald's QCP) for the largest
// eigenvalue, eigenvector from the adjugate of K - lambda I, then ONE first-order correction in the frame
// rotated by the approximate answer (which squares the error, so the polynomial root only has to be good to
// ~1e-3).  ~400 instructions instead of ~1300 for converged Jacobi.  Returns false when the correction was
// not small (near-degenerate top eigenvalue, slow Newton) -- the caller then runs Jacobi for that frame,
// so accuracy never rests on the fast path alone.  `S` is the covariance scaled to unit Frobenius norm.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_to_rot(const float (&q)[4], float (&R)[9]) {
  const float qa = q[0], qb = q[1], qc = q[2], qd = q[3];
  const float aa = qa * qa, bb = qb * qb, cc = qc * qc, dd = qd * qd;
  const float bc = qb * qc, ad = qa * qd, bd = qb * qd, ac = qa * qc, cd = qc * qd, ab = qa * qb;
  R[0] = (aa + bb) - (cc + dd);
  R[1] = 2.f * (bc + ad);
  R[2] = 2.f * (bd - ac);
  R[3] = 2.f * (bc - ad);
  R[4] = (aa - bb) + (cc - dd);
  R[5] = 2.f * (cd + ab);
  R[6] = 2.f * (bd + ac);
  R[7] = 2.f * (cd - ab);
  R[8] = (aa - bb) - (cc - dd);
}

__device__ __forceinline__ bool dominant_quat_fast(const float (&S)[9], float (&q)[4]) {
  const float Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
  // Horn's matrix (symmetric, trace 0)
  const float k00 = Sxx + Syy + Szz, k01 = Syz - Szy, k02 = Szx - Sxz, k03 = Sxy - Syx;
  const float k11 = Sxx - Syy - Szz, k12 = Sxy + Syx, k13 = Szx + Sxz;
  const float k22 = -Sxx + Syy - Szz, k23 = Syz + Szy;
  const float k33 = -Sxx - Syy + Szz;
  // det(K - lambda I) = l^4 + c2 l^2 + c1 l + c0
  float ss = Sxx * Sxx;
  ss = fmaf(Sxy, Sxy, ss); ss = fmaf(Sxz, Sxz, ss); ss = fmaf(Syx, Syx, ss); ss = fmaf(Syy, Syy, ss);
  ss = fmaf(Syz, Syz, ss); ss = fmaf(Szx, Szx, ss); ss = fmaf(Szy, Szy, ss); ss = fmaf(Szz, Szz, ss);
  const float c2 = -2.0f * ss;
  const float detS = Sxx * (Syy * Szz - Syz * Szy) - Sxy * (Syx * Szz - Syz * Szx) + Sxz * (Syx * Szy - Syy * Szx);
  const float c1 = -8.0f * detS;
  float c0;
  {
    const float s0 = k00 * k11 - k01 * k01, s1 = k00 * k12 - k01 * k02, s2 = k00 * k13 - k01 * k03;
    const float s3 = k01 * k12 - k11 * k02, s4 = k01 * k13 - k11 * k03, s5 = k02 * k13 - k12 * k03;
    const float m5 = k22 * k33 - k23 * k23, m4 = k12 * k33 - k13 * k23, m3 = k12 * k23 - k13 * k22;
    const float m2 = k02 * k33 - k03 * k23, m1 = k02 * k23 - k03 * k22, m0 = k02 * k13 - k03 * k12;
    c0 = s0 * m5 - s1 * m4 + s2 * m3 + s3 * m2 - s4 * m1 + s5 * m0;
  }
  // Newton from the upper bound sqrt(3) ||S||_F (monotone from above: all roots are real)
  float lam = 1.7320508f;
#pragma unroll 1
  for (int it = 0; it < 6; ++it) {          // rolled: the fused kernels are instruction-fetch sensitive
    const float l2 = lam * lam;
    const float P = fmaf(fmaf(l2 + c2, lam, c1), lam, c0);
    const float dP = fmaf(fmaf(4.0f, l2, 2.0f * c2), lam, c1);
    lam -= __fdividef(P, dP);
  }
  // adjugate of B = K - lambda I (symmetric); its columns are all proportional to the eigenvector
  const float b00 = k00 - lam, b11 = k11 - lam, b22 = k22 - lam, b33 = k33 - lam;
  const float s0 = b00 * b11 - k01 * k01, s1 = b00 * k12 - k01 * k02, s2 = b00 * k13 - k01 * k03;
  const float s3 = k01 * k12 - b11 * k02, s4 = k01 * k13 - b11 * k03, s5 = k02 * k13 - k12 * k03;
  const float m5 = b22 * b33 - k23 * k23, m4 = k12 * b33 - k13 * k23, m3 = k12 * k23 - k13 * b22;
  const float m2 = k02 * b33 - k03 * k23, m1 = k02 * k23 - k03 * b22, m0 = k02 * k13 - k03 * k12;
  const float a00 = b11 * m5 - k12 * m4 + k13 * m3;
  const float a01 = -k01 * m5 + k02 * m4 - k03 * m3;
  const float a02 = k13 * s5 - k23 * s4 + b33 * s3;
  const float a03 = -k12 * s5 + b22 * s4 - k23 * s3;
  const float a11 = b00 * m5 - k02 * m2 + k03 * m1;
  const float a12 = -k03 * s5 + k23 * s2 - b33 * s1;
  const float a13 = k02 * s5 - b22 * s2 + k23 * s1;
  const float a22 = k03 * s4 - k13 * s2 + b33 * s0;
  const float a23 = -k02 * s4 + k12 * s2 - k23 * s0;
  const float a33 = k02 * s3 - k12 * s1 + b22 * s0;
  // column with the largest diagonal entry (|a_ii| ~ q_i^2 >= 1/4 of the trace)
  float q0[4] = {a00, a01, a02, a03};
  float best = fabsf(a00);
  if (fabsf(a11) > best) { best = fabsf(a11); q0[0] = a01; q0[1] = a11; q0[2] = a12; q0[3] = a13; }
  if (fabsf(a22) > best) { best = fabsf(a22); q0[0] = a02; q0[1] = a12; q0[2] = a22; q0[3] = a23; }
  if (fabsf(a33) > best) { best = fabsf(a33); q0[0] = a03; q0[1] = a13; q0[2] = a23; q0[3] = a33; }
  {
    const float n2 = fmaf(q0[0], q0[0], fmaf(q0[1], q0[1], fmaf(q0[2], q0[2], q0[3] * q0[3])));
    const float inv = rsqrtf(n2);
#pragma unroll
    for (int r = 0; r < 4; ++r) q0[r] *= inv;
  }
  // first-order correction: with S' = R0^T S the residual rotation vector d solves (tr(P) I - P) d = vee(S')/2,
  // P = sym(S'); the corrected quaternion is (1, d) * q0.
  float R0[9];
  quat_to_rot(q0, R0);
  float T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) T[3 * i + k] = fmaf(R0[i], S[k], fmaf(R0[3 + i], S[3 + k], R0[6 + i] * S[6 + k]));
  const float p01 = 0.5f * (T[1] + T[3]), p02 = 0.5f * (T[2] + T[6]), p12 = 0.5f * (T[5] + T[7]);
  const float e00 = T[4] + T[8], e11 = T[0] + T[8], e22 = T[0] + T[4];     // tr(P) - P_ii
  const float r0 = 0.5f * (T[5] - T[7]), r1 = 0.5f * (T[6] - T[2]), r2 = 0.5f * (T[1] - T[3]);
  const float g00 = e11 * e22 - p12 * p12;
  const float g01 = p01 * e22 + p02 * p12;         // cofactors of [[e00,-p01,-p02],[-p01,e11,-p12],[-p02,-p12,e22]]
  const float g02 = p01 * p12 + p02 * e11;
  const float g11 = e00 * e22 - p02 * p02;
  const float g12 = e00 * p12 + p01 * p02;
  const float g22 = e00 * e11 - p01 * p01;
  const float det = e00 * g00 - p01 * g01 - p02 * g02;
  const float idet = __fdividef(1.0f, det);
  const float d0 = (g00 * r0 + g01 * r1 + g02 * r2) * idet;
  const float d1 = (g01 * r0 + g11 * r1 + g12 * r2) * idet;
  const float d2 = (g02 * r0 + g12 * r1 + g22 * r2) * idet;
  const float dn2 = fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
  // (1, d) * (a, v) = (a - d.v,  v + a d + d x v)
  const float a = q0[0], vx = q0[1], vy = q0[2], vz = q0[3];
  float qa = a - (d0 * vx + d1 * vy + d2 * vz);
  float qx = vx + a * d0 + (d1 * vz - d2 * vy);
  float qy = vy + a * d1 + (d2 * vx - d0 * vz);
  float qz = vz + a * d2 + (d0 * vy - d1 * vx);
  const float inv = rsqrtf(fmaf(qa, qa, fmaf(qx, qx, fmaf(qy, qy, qz * qz))));
  q[0] = qa * inv; q[1] = qx * inv; q[2] = qy * inv; q[3] = qz * inv;
  return dn2 < 1e-6f;             // false also for NaN
}

struct Rigid {
  float c[3];   // centroid of the alignment selection
  float R[9];   // row-major rotation, z = (x - c) R
  float H[9];   // row-major covariance  sum_k (x_k - c)^T Y_k
};

// Optimal superposition of frame `xf` ([n,3]) onto the centred reference.  All lanes of the group
// return the same result.  aidx / refx may live in shared or global memory.
//
// One pass over the selection: sums are taken relative to a pivot atom p = x[A_0] (d_k = x_k - p), which
// keeps fp32 accurate for frames far from the origin; since the reference is centred (sum_k y_k = 0),
//   H = sum_k (x_k - c)^T y_k = sum_k d_k^T y_k      and      c = p + mean_k d_k.
template <int G>
__device__ __forceinline__ void kabsch_moments(const float* __restrict__ xf, const int* __restrict__ aidx,
                                               const float* __restrict__ refx, int n_align, int lane, Rigid& rg) {
  const float* p0 = xf + 3 * aidx[0];
  const float pvx = p0[0], pvy = p0[1], pvz = p0[2];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  float h[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = 0.f;
  for (int k = lane; k < n_align; k += G) {
    const float* p = xf + 3 * aidx[k];
    const float px = p[0] - pvx, py = p[1] - pvy, pz = p[2] - pvz;
    const float y0 = refx[3 * k], y1 = refx[3 * k + 1], y2 = refx[3 * k + 2];
    sx += px; sy += py; sz += pz;
    h[0] = fmaf(px, y0, h[0]); h[1] = fmaf(px, y1, h[1]); h[2] = fmaf(px, y2, h[2]);
    h[3] = fmaf(py, y0, h[3]); h[4] = fmaf(py, y1, h[4]); h[5] = fmaf(py, y2, h[5]);
    h[6] = fmaf(pz, y0, h[6]); h[7] = fmaf(pz, y1, h[7]); h[8] = fmaf(pz, y2, h[8]);
  }
  const float inv_n = 1.0f / (float)n_align;
  rg.c[0] = fmaf(gsum<G>(sx), inv_n, pvx);
  rg.c[1] = fmaf(gsum<G>(sy), inv_n, pvy);
  rg.c[2] = fmaf(gsum<G>(sz), inv_n, pvz);
#pragma unroll
  for (int i = 0; i < 9; ++i) rg.H[i] = gsum<G>(h[i]);
}

struct Quat4 {
  float a, b, c, d;
};
// Jacobi route for the frames the fast path declined (kept out of line: it is rare, and the hot path of the
// fused kernels should stay short and contiguous for the instruction cache).
#ifndef MOLANN_HOST_EMULATION
__device__ __noinline__
#else
inline
#endif
Quat4 kabsch_rotation_jacobi(float Sxx, float Sxy, float Sxz, float Syx, float Syy, float Syz, float Szx, float Szy,
                             float Szz, bool need) {
  float a[4][4];
  a[0][0] = Sxx + Syy + Szz; a[0][1] = Syz - Szy; a[0][2] = Szx - Sxz; a[0][3] = Sxy - Syx;
  a[1][1] = Sxx - Syy - Szz; a[1][2] = Sxy + Syx; a[1][3] = Szx + Sxz;
  a[2][2] = -Sxx + Syy - Szz; a[2][3] = Syz + Szy;
  a[3][3] = -Sxx - Syy + Szz;
  a[1][0] = a[2][0] = a[2][1] = a[3][0] = a[3][1] = a[3][2] = 0.f;   // lower triangle unused
  float qj[4];
  dominant_eigvec4(a, qj, need);
  Quat4 r;
  r.a = qj[0]; r.b = qj[1]; r.c = qj[2]; r.d = qj[3];
  return r;
}

// rg.R from rg.H.  Must be called by all 32 lanes of a warp (the Jacobi fallback votes).
__device__ __forceinline__ void kabsch_rotation(Rigid& rg) {
  float nrm2 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) nrm2 = fmaf(rg.H[i], rg.H[i], nrm2);
  const float inv = rsqrtf(nrm2);
  float S[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) S[i] = rg.H[i] * inv;
  float q[4];
  const bool ok = dominant_quat_fast(S, q);
  if (!__all_sync(0xffffffffu, ok)) {          // rare: near-degenerate frames (arguments by value: no stack traffic)
    const Quat4 qj = kabsch_rotation_jacobi(rg.H[0], rg.H[1], rg.H[2], rg.H[3], rg.H[4], rg.H[5], rg.H[6], rg.H[7],
                                            rg.H[8], !ok);
    if (!ok) { q[0] = qj.a; q[1] = qj.b; q[2] = qj.c; q[3] = qj.d; }
  }
  quat_to_rot(q, rg.R);
}

template <int G>
__device__ __forceinline__ void kabsch(const float* __restrict__ xf, const int* __restrict__ aidx,
                                       const float* __restrict__ refx, int n_align, int lane, Rigid& rg) {
  kabsch_moments<G>(xf, aidx, refx, n_align, lane, rg);
  kabsch_rotation(rg);
}

// z = (p - c) R
__device__ __forceinline__ void rigid_apply(const Rigid& rg, float px, float py, float pz, float& zx,
                                            float& zy, float& zz) {
  const float dx = px - rg.c[0], dy = py - rg.c[1], dz = pz - rg.c[2];
  zx = fmaf(dx, rg.R[0], fmaf(dy, rg.R[3], dz * rg.R[6]));
  zy = fmaf(dx, rg.R[1], fmaf(dy, rg.R[4], dz * rg.R[7]));
  zz = fmaf(dx, rg.R[2], fmaf(dy, rg.R[5], dz * rg.R[8]));
}

// t = g R^T
__device__ __forceinline__ void rot_transpose_apply(const Rigid& rg, float gx, float gy, float gz, float& tx,
                                                    float& ty, float& tz) {
  tx = fmaf(gx, rg.R[0], fmaf(gy, rg.R[1], gz * rg.R[2]));
  ty = fmaf(gx, rg.R[3], fmaf(gy, rg.R[4], gz * rg.R[5]));
  tz = fmaf(gx, rg.R[6], fmaf(gy, rg.R[7], gz * rg.R[8]));
}

struct Entry {
  int type, a0, a1, a2, a3, off;
};
__device__ __forceinline__ Entry load_entry(const int* __restrict__ e) {
  Entry r;
  r.type = e[0]; r.a0 = e[1]; r.a1 = e[2]; r.a2 = e[3]; r.a3 = e[4]; r.off = e[5];
  return r;
}

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 ld3(const float* __restrict__ xf, int a) {
  const float* p = xf + 3 * a;
  V3 v; v.x = p[0]; v.y = p[1]; v.z = p[2];
  return v;
}
__device__ __forceinline__ V3 sub(V3 a, V3 b) { V3 r; r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z; return r; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { V3 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; return r; }
__device__ __forceinline__ V3 scale(V3 a, float s) { V3 r; r.x = a.x * s; r.y = a.y * s; r.z = a.z * s; return r; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  V3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}

// ------------------------------------------------------------------------------------------------
// Feature forward for one program entry.  `out(col, value)` stores one feature column.
// ------------------------------------------------------------------------------------------------
template <class Out>
__device__ __forceinline__ void feature_forward(const Entry& e, const float* __restrict__ xf, bool aligned,
                                                const Rigid& rg, int use_angle, Out& out) {
  if (e.type == FEAT_POSITION) {
    const V3 p = ld3(xf, e.a0);
    float zx = p.x, zy = p.y, zz = p.z;
    if (aligned) rigid_apply(rg, p.x, p.y, p.z, zx, zy, zz);
    out(e.off, zx); out(e.off + 1, zy); out(e.off + 2, zz);
  } else if (e.type == FEAT_DIHEDRAL) {        // reference ann.py:338-351
    const V3 x0 = ld3(xf, e.a0), x1 = ld3(xf, e.a1), x2 = ld3(xf, e.a2), x3 = ld3(xf, e.a3);
    const V3 r12 = sub(x1, x0), r23 = sub(x2, x1), r34 = sub(x3, x2);
    const V3 n1 = cross(r12, r23), n2 = cross(r23, r34);
    const float C = dot(n1, n2);
    const float S = dot(n1, r34) * sqrtf(dot(r23, r23));
    if (use_angle) {
      out(e.off, atan2f(S, C));
    } else {
      const float rho = sqrtf(fmaf(C, C, S * S));
      out(e.off, C / rho); out(e.off + 1, S / rho);
    }
  } else if (e.type == FEAT_BOND) {            // ann.py:334-336
    const V3 r = sub(ld3(xf, e.a1), ld3(xf, e.a0));
    out(e.off, sqrtf(dot(r, r)));
  } else {                                     // angle, ann.py:323-332
    const V3 xv = ld3(xf, e.a1);
    const V3 u = sub(ld3(xf, e.a0), xv), w = sub(ld3(xf, e.a2), xv);
    const float cs = dot(u, w) / (sqrtf(dot(u, u)) * sqrtf(dot(w, w)));
    out(e.off, use_angle ? acosf(cs) : cs);
  }
}

// Four atoms of one entry at once; accumulators that can do better than four separate calls (vector reductions
// for four consecutive atoms, staged_block.cuh) overload this.
template <class Acc>
__device__ __forceinline__ void acc_quad(Acc& acc, int a0, int a1, int a2, int a3, V3 g0, V3 g1, V3 g2, V3 g3) {
  acc(a0, g0); acc(a1, g1); acc(a2, g2); acc(a3, g3);
}

// ------------------------------------------------------------------------------------------------
// Feature backward for one entry.  `gin(col)` reads the cotangent of a feature column,
// `acc(atom, V3)` accumulates into the frame's coordinate gradient.  For aligned position entries the
// 3x3 moment M = (x-c)^T G and sg = sum G R^T needed by the alignment backward are accumulated too.
// ------------------------------------------------------------------------------------------------
template <class GIn, class Acc>
__device__ __forceinline__ void feature_backward(const Entry& e, const float* __restrict__ xf, bool aligned,
                                                 const Rigid& rg, int use_angle, GIn& gin, Acc& acc,
                                                 float (&M)[9], float (&sg)[3]) {
  if (e.type == FEAT_POSITION) {
    float g0, g1, g2;
    gin.load3(e.off, g0, g1, g2);
    if (aligned) {
      const V3 p = ld3(xf, e.a0);
      const float dx = p.x - rg.c[0], dy = p.y - rg.c[1], dz = p.z - rg.c[2];
      M[0] = fmaf(dx, g0, M[0]); M[1] = fmaf(dx, g1, M[1]); M[2] = fmaf(dx, g2, M[2]);
      M[3] = fmaf(dy, g0, M[3]); M[4] = fmaf(dy, g1, M[4]); M[5] = fmaf(dy, g2, M[5]);
      M[6] = fmaf(dz, g0, M[6]); M[7] = fmaf(dz, g1, M[7]); M[8] = fmaf(dz, g2, M[8]);
      V3 t;
      rot_transpose_apply(rg, g0, g1, g2, t.x, t.y, t.z);
      sg[0] += t.x; sg[1] += t.y; sg[2] += t.z;
      acc(e.a0, t);
    } else {
      V3 t; t.x = g0; t.y = g1; t.z = g2;
      acc(e.a0, t);
    }
  } else if (e.type == FEAT_DIHEDRAL) {
    const V3 x0 = ld3(xf, e.a0), x1 = ld3(xf, e.a1), x2 = ld3(xf, e.a2), x3 = ld3(xf, e.a3);
    const V3 r12 = sub(x1, x0), r23 = sub(x2, x1), r34 = sub(x3, x2);
    const V3 n1 = cross(r12, r23), n2 = cross(r23, r34);
    const float l23 = sqrtf(dot(r23, r23));
    const float d1 = dot(n1, r34);
    const float C = dot(n1, n2);
    const float S = d1 * l23;
    const float rho2 = fmaf(C, C, S * S);
    float gC, gS;
    if (use_angle) {
      const float g = gin(e.off);
      gC = -g * S / rho2;
      gS = g * C / rho2;
    } else {
      float gc, gs;
      gin.load2(e.off, gc, gs);
      const float rho3 = rho2 * sqrtf(rho2);
      const float k = (gc * S - gs * C) / rho3;
      gC = S * k;
      gS = -C * k;
    }
    // C = n1.n2 ; S = (n1.r34) * |r23|
    V3 gn1 = add(scale(n2, gC), scale(r34, gS * l23));
    V3 gn2 = scale(n1, gC);
    V3 gr34 = scale(n1, gS * l23);
    V3 gr23 = scale(r23, gS * d1 / l23);
    // n1 = r12 x r23 ; n2 = r23 x r34     (c = a x b: g_a = b x g_c, g_b = g_c x a)
    V3 gr12 = cross(r23, gn1);
    gr23 = add(gr23, cross(gn1, r12));
    gr23 = add(gr23, cross(r34, gn2));
    gr34 = add(gr34, cross(gn2, r23));
    acc_quad(acc, e.a0, e.a1, e.a2, e.a3, scale(gr12, -1.f), sub(gr12, gr23), sub(gr23, gr34), gr34);
  } else if (e.type == FEAT_BOND) {
    const V3 r = sub(ld3(xf, e.a1), ld3(xf, e.a0));
    const float g = gin(e.off) / sqrtf(dot(r, r));
    acc(e.a1, scale(r, g));
    acc(e.a0, scale(r, -g));
  } else {  // angle
    const V3 xv = ld3(xf, e.a1);
    const V3 u = sub(ld3(xf, e.a0), xv), w = sub(ld3(xf, e.a2), xv);
    const float lu = sqrtf(dot(u, u)), lw = sqrtf(dot(w, w));
    const float inv = 1.0f / (lu * lw);
    const float cs = dot(u, w) * inv;
    float g = gin(e.off);
    if (use_angle) g = -g / sqrtf(1.0f - cs * cs);
    const V3 gu = scale(sub(scale(w, inv), scale(u, cs / (lu * lu))), g);
    const V3 gw = scale(sub(scale(u, inv), scale(w, cs / (lw * lw))), g);
    acc(e.a0, gu);
    acc(e.a2, gw);
    acc(e.a1, scale(add(gu, gw), -1.f));
  }
}

// dL/dH (row-major 3x3) from the reduced moment M and the frame's rigid transform.
__device__ __forceinline__ void align_backward_dH(const Rigid& rg, const float (&M)[9], float (&dH)[9]) {
  const float* R = rg.R;
  const float* H = rg.H;
  // P = R^T H (symmetric in exact arithmetic; symmetrised here)
  float P[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) P[3 * i + j] = R[i] * H[j] + R[3 + i] * H[3 + j] + R[6 + i] * H[6 + j];
  const float p01 = 0.5f * (P[1] + P[3]), p02 = 0.5f * (P[2] + P[6]), p12 = 0.5f * (P[5] + P[7]);
  const float tr = P[0] + P[4] + P[8];
  // B = tr(P) I - P
  const float b00 = tr - P[0], b11 = tr - P[4], b22 = tr - P[8];
  const float b01 = -p01, b02 = -p02, b12 = -p12;
  // A = R^T M - M^T R  (antisymmetric); ax = (A21, A02, A10)
  float RtM[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) RtM[3 * i + j] = R[i] * M[j] + R[3 + i] * M[3 + j] + R[6 + i] * M[6 + j];
  const float ax0 = RtM[7] - RtM[5];
  const float ax1 = RtM[2] - RtM[6];
  const float ax2 = RtM[3] - RtM[1];
  // solve B w = ax by the adjugate (B symmetric)
  const float c00 = b11 * b22 - b12 * b12;
  const float c01 = b02 * b12 - b01 * b22;
  const float c02 = b01 * b12 - b02 * b11;
  const float c11 = b00 * b22 - b02 * b02;
  const float c12 = b01 * b02 - b00 * b12;
  const float c22 = b00 * b11 - b01 * b01;
  const float det = b00 * c00 + b01 * c01 + b02 * c02;
  const float idet = 1.0f / det;
  const float w0 = (c00 * ax0 + c01 * ax1 + c02 * ax2) * idet;
  const float w1 = (c01 * ax0 + c11 * ax1 + c12 * ax2) * idet;
  const float w2 = (c02 * ax0 + c12 * ax1 + c22 * ax2) * idet;
  // dH = R [w]x,  [w]x = [[0,-w2,w1],[w2,0,-w0],[-w1,w0,0]]
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float r0 = R[3 * i], r1 = R[3 * i + 1], r2 = R[3 * i + 2];
    dH[3 * i + 0] = r1 * w2 - r2 * w1;
    dH[3 * i + 1] = r2 * w0 - r0 * w2;
    dH[3 * i + 2] = r0 * w1 - r1 * w0;
  }
}

// gradient contribution of alignment atom k (reference row y): y dH^T - sg / n_a
__device__ __forceinline__ V3 align_atom_grad(const float (&dH)[9], const float (&sg)[3], float inv_na, float y0,
                                              float y1, float y2) {
  V3 t;
  t.x = fmaf(y0, dH[0], fmaf(y1, dH[1], y2 * dH[2])) - sg[0] * inv_na;
  t.y = fmaf(y0, dH[3], fmaf(y1, dH[4], y2 * dH[5])) - sg[1] * inv_na;
  t.z = fmaf(y0, dH[6], fmaf(y1, dH[7], y2 * dH[8])) - sg[2] * inv_na;
  return t;
}

__device__ __forceinline__ float act_forward(float v, int act) {
  switch (act) {
    case ACT_TANH: return tanhf(v);
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}
// derivative expressed through the activation OUTPUT h
__device__ __forceinline__ float act_grad_from_output(float h, int act) {
  switch (act) {
    case ACT_TANH: return fmaf(-h, h, 1.0f);
    case ACT_RELU: return h > 0.f ? 1.0f : 0.f;
    case ACT_SIGMOID: return h * (1.0f - h);
    default: return 1.0f;
  }
}

}  // namespace molann
