// host_emulation.cpp -- TEST SCAFFOLDING: runs the device geometry functions of
// molann_b200/csrc/geometry.cuh on the CPU (one "thread" per frame, G = 1) so pytest can compare the
// closed-form forward/backward math with the oracle in the GPU-less container.  Not part of the product.
#define MOLANN_HOST_EMULATION 1
#include "../../molann_b200/csrc/geometry.cuh"

#include <cstring>
#include <vector>

using namespace molann;

namespace {
struct RowOut { float* row; void operator()(int col, float v) { row[col] = v; } };
struct RowGIn {
  const float* row;
  float operator()(int col) const { return row[col]; }
  void load2(int col, float& a, float& b) const { a = row[col]; b = row[col + 1]; }
  void load3(int col, float& a, float& b, float& c) const { a = row[col]; b = row[col + 1]; c = row[col + 2]; }
};
struct RowAccH { float* row; void operator()(int a, V3 v) { row[3*a] += v.x; row[3*a+1] += v.y; row[3*a+2] += v.z; } };
}

extern "C" {

void emu_preprocess_forward(int n_inp, int n_align, const int* aidx, const float* ref, int n_entries,
                            const int* entries, int d_feat, int use_angle, const float* x, long long L, float* feat) {
  for (long long f = 0; f < L; ++f) {
    const float* xf = x + f * 3 * n_inp;
    Rigid rg;
    const bool aligned = n_align > 0;
    if (aligned) kabsch<1>(xf, aidx, ref, n_align, 0, rg);
    RowOut out{feat + f * d_feat};
    for (int e = 0; e < n_entries; ++e) {
      const Entry en = load_entry(entries + ENTRY_INTS * e);
      feature_forward(en, xf, aligned, rg, use_angle, out);
    }
  }
}

void emu_preprocess_backward(int n_inp, int n_align, const int* aidx, const float* ref, int n_entries,
                             const int* entries, int d_feat, int use_angle, const float* x, const float* gfeat,
                             long long L, float* gx) {
  for (long long f = 0; f < L; ++f) {
    const float* xf = x + f * 3 * n_inp;
    float* gxf = gx + f * 3 * n_inp;
    std::memset(gxf, 0, sizeof(float) * 3 * n_inp);
    Rigid rg;
    const bool aligned = n_align > 0;
    if (aligned) kabsch<1>(xf, aidx, ref, n_align, 0, rg);
    RowGIn gin{gfeat + f * d_feat};
    RowAccH acc{gxf};
    float M[9] = {0}, sg[3] = {0};
    for (int e = 0; e < n_entries; ++e) {
      const Entry en = load_entry(entries + ENTRY_INTS * e);
      feature_backward(en, xf, aligned, rg, use_angle, gin, acc, M, sg);
    }
    if (aligned) {
      float dH[9];
      align_backward_dH(rg, M, dH);
      const float inv_na = 1.0f / (float)n_align;
      for (int k = 0; k < n_align; ++k)
        acc(aidx[k], align_atom_grad(dH, sg, inv_na, ref[3*k], ref[3*k+1], ref[3*k+2]));
    }
  }
}

void emu_align_forward(int n_inp, int n_align, const int* aidx, const float* ref, const float* x, long long L,
                       float* out) {
  for (long long f = 0; f < L; ++f) {
    const float* xf = x + f * 3 * n_inp;
    Rigid rg;
    kabsch<1>(xf, aidx, ref, n_align, 0, rg);
    for (int j = 0; j < n_inp; ++j)
      rigid_apply(rg, xf[3*j], xf[3*j+1], xf[3*j+2], out[f*3*n_inp + 3*j], out[f*3*n_inp + 3*j+1], out[f*3*n_inp + 3*j+2]);
  }
}

void emu_align_backward(int n_inp, int n_align, const int* aidx, const float* ref, const float* x,
                        const float* gout, long long L, float* gx) {
  for (long long f = 0; f < L; ++f) {
    const float* xf = x + f * 3 * n_inp;
    const float* gf = gout + f * 3 * n_inp;
    float* gxf = gx + f * 3 * n_inp;
    Rigid rg;
    kabsch<1>(xf, aidx, ref, n_align, 0, rg);
    float M[9] = {0}, sg[3] = {0};
    for (int j = 0; j < n_inp; ++j) {
      const float g0 = gf[3*j], g1 = gf[3*j+1], g2 = gf[3*j+2];
      const float dx = xf[3*j] - rg.c[0], dy = xf[3*j+1] - rg.c[1], dz = xf[3*j+2] - rg.c[2];
      M[0] += dx*g0; M[1] += dx*g1; M[2] += dx*g2; M[3] += dy*g0; M[4] += dy*g1; M[5] += dy*g2;
      M[6] += dz*g0; M[7] += dz*g1; M[8] += dz*g2;
      float tx, ty, tz;
      rot_transpose_apply(rg, g0, g1, g2, tx, ty, tz);
      sg[0] += tx; sg[1] += ty; sg[2] += tz;
      gxf[3*j] = tx; gxf[3*j+1] = ty; gxf[3*j+2] = tz;
    }
    float dH[9];
    align_backward_dH(rg, M, dH);
    RowAccH acc{gxf};
    const float inv_na = 1.0f / (float)n_align;
    for (int k = 0; k < n_align; ++k)
      acc(aidx[k], align_atom_grad(dH, sg, inv_na, ref[3*k], ref[3*k+1], ref[3*k+2]));
  }
}

// rotation of every frame (row-major R, z = (x - c) R) and whether the fast polynomial path accepted it
void emu_kabsch(int n_inp, int n_align, const int* aidx, const float* ref, const float* x, long long L, float* Rout,
                int* fast_ok) {
  for (long long f = 0; f < L; ++f) {
    const float* xf = x + f * 3 * n_inp;
    Rigid rg;
    kabsch<1>(xf, aidx, ref, n_align, 0, rg);
    for (int i = 0; i < 9; ++i) Rout[9 * f + i] = rg.R[i];
    float n2 = 0.f, S[9], q[4];
    for (int i = 0; i < 9; ++i) n2 += rg.H[i] * rg.H[i];
    for (int i = 0; i < 9; ++i) S[i] = rg.H[i] / sqrtf(n2);
    fast_ok[f] = dominant_quat_fast(S, q) ? 1 : 0;
  }
}

float emu_act_forward(float v, int act) { return act_forward(v, act); }
float emu_act_grad(float h, int act) { return act_grad_from_output(h, act); }

}  // extern "C"
