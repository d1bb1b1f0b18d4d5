// sb_trace.cu -- DEVELOPMENT TOOL: phase timeline (clock64, thread 0 of CTA 0) of the block-per-frame preprocess
// backward kernel on a C3-shaped problem (2000 atoms, every 10th aligned + position feature, 100 dihedrals).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMOLANN_WS_TRACE -I../../include -o sb_trace sb_trace.cu
#include "../../molann_b200/csrc/molann_b200.cu"

#include <cstdio>
#include <vector>

int main() {
  const int n = 2000, na = 200, L = 8192;
  std::vector<float> x((size_t)L * n * 3), ref(na * 3), base(n * 3);
  unsigned s = 777u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  for (int j = 0; j < n; ++j) { base[3 * j] = 1.5f * j * 0.3f + rnd(); base[3 * j + 1] = 10.f * rnd(); base[3 * j + 2] = 10.f * rnd(); }
  for (size_t f = 0; f < (size_t)L; ++f)
    for (int j = 0; j < n * 3; ++j) x[f * n * 3 + j] = base[j] + 0.2f * rnd();
  std::vector<int> aidx(na), ent;
  float c[3] = {0, 0, 0};
  for (int k = 0; k < na; ++k) { aidx[k] = 10 * k; for (int d = 0; d < 3; ++d) c[d] += base[30 * k + d] / na; }
  for (int k = 0; k < na; ++k) for (int d = 0; d < 3; ++d) ref[3 * k + d] = base[30 * k + d] - c[d];
  int col = 0;
  for (int k = 0; k < na; ++k) { int e[6] = {3, 10 * k, 0, 0, 0, col}; ent.insert(ent.end(), e, e + 6); col += 3; }
  for (int k = 0; k < 100; ++k) { int e[6] = {2, 20 * k, 20 * k + 1, 20 * k + 2, 20 * k + 3, col}; ent.insert(ent.end(), e, e + 6); col += 2; }
  float *dx, *dgf, *dgx, *dref; int *daidx, *dent;
  cudaMalloc(&dx, x.size() * 4); cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dgx, x.size() * 4);
  std::vector<float> gf((size_t)L * col);
  for (auto& v : gf) v = rnd();
  cudaMalloc(&dgf, gf.size() * 4); cudaMemcpy(dgf, gf.data(), gf.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dref, ref.size() * 4); cudaMemcpy(dref, ref.data(), ref.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&daidx, na * 4); cudaMemcpy(daidx, aidx.data(), na * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dent, ent.size() * 4); cudaMemcpy(dent, ent.data(), ent.size() * 4, cudaMemcpyHostToDevice);
  MolannPlan p; std::memset(&p, 0, sizeof(p));
  p.n_inp = n; p.n_align = na; p.align_idx = daidx; p.ref_x = dref; p.n_entries = (int)ent.size() / 6; p.entries = dent;
  p.d_feat = col; p.use_angle_value = 0; p.n_layers = 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int st = molann_b200_preprocess_backward(&p, dx, dgf, L, dgx, nullptr);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("# preprocess_backward status %d  %.3f ms for %d frames (%.1f M frames/s)\n", st, ms, L, L / ms * 1e-3);
  }
  static long long tr[64 * 16];
  cudaMemcpyFromSymbol(tr, molann::g_sb_trace, sizeof(tr));
  const char* ev[11] = {"top", "zero", "land+moments(next)+sync", "invariant", "wait-rigid", "positions+red", "sync", "dH", "sync", "scatter", "sync"};
  for (int i = 2; i < 8; ++i) {
    printf("frame %2d:", i);
    for (int e = 1; e < 11; ++e) printf(" %s +%lld", ev[e], tr[i * 16 + e] - tr[i * 16 + e - 1]);
    printf(" | total %lld\n", tr[(i + 1) * 16] - tr[i * 16]);
  }
  return 0;
}
