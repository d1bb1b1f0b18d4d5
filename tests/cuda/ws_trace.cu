// ws_trace.cu -- DEVELOPMENT TOOL (not part of the product, not used by the tests): builds the kernels with
// MOLANN_WS_TRACE, runs the warp-specialised forward kernel on a C2-shaped problem and prints the clock64()
// stamps CTA 0 recorded for every role and tile, so pipeline bubbles can be read off a timeline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMOLANN_WS_TRACE -o ws_trace ws_trace.cu && ./ws_trace
#include "../../molann_b200/csrc/molann_b200.cu"

#include <cstdio>
#include <vector>

int main() {
  const int n_inp = 22, n_align = 10, L = 148 * 128 * 40;
  const int heavy[10] = {1, 4, 5, 6, 8, 10, 14, 15, 16, 18};
  std::vector<float> x((size_t)L * n_inp * 3), ref(30);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  std::vector<float> base(n_inp * 3);
  for (auto& v : base) v = 6.0f * rnd();
  for (size_t f = 0; f < (size_t)L; ++f)
    for (int j = 0; j < n_inp * 3; ++j) x[f * n_inp * 3 + j] = base[j] + 0.3f * rnd();
  float c[3] = {0, 0, 0};
  for (int k = 0; k < 10; ++k) for (int d = 0; d < 3; ++d) c[d] += base[3 * heavy[k] + d] / 10.f;
  for (int k = 0; k < 10; ++k) for (int d = 0; d < 3; ++d) ref[3 * k + d] = base[3 * heavy[k] + d] - c[d];
  std::vector<int> ent(10 * 6);
  for (int k = 0; k < 10; ++k) { int* e = &ent[6 * k]; e[0] = 3; e[1] = heavy[k]; e[2] = e[3] = e[4] = 0; e[5] = 3 * k; }
  const int dims[4] = {30, 64, 64, 2};
  float *dx, *dy, *dref, *dW[3], *db[3]; int *daidx, *dent;
  cudaMalloc(&dx, x.size() * 4); cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dy, (size_t)L * 2 * 4);
  cudaMalloc(&dref, 120); cudaMemcpy(dref, ref.data(), 120, cudaMemcpyHostToDevice);
  cudaMalloc(&daidx, 40); cudaMemcpy(daidx, heavy, 40, cudaMemcpyHostToDevice);
  cudaMalloc(&dent, ent.size() * 4); cudaMemcpy(dent, ent.data(), ent.size() * 4, cudaMemcpyHostToDevice);
  MolannPlan p; std::memset(&p, 0, sizeof(p));
  p.n_inp = n_inp; p.n_align = n_align; p.align_idx = daidx; p.ref_x = dref; p.n_entries = 10; p.entries = dent;
  p.d_feat = 30; p.use_angle_value = 0; p.n_layers = 3; p.act_id = 0;
  for (int k = 0; k < 4; ++k) p.dims[k] = dims[k];
  for (int k = 0; k < 3; ++k) {
    std::vector<float> w(dims[k] * dims[k + 1]), b(dims[k + 1]);
    for (auto& v : w) v = 0.3f * rnd();
    for (auto& v : b) v = 0.1f * rnd();
    cudaMalloc(&dW[k], w.size() * 4); cudaMemcpy(dW[k], w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&db[k], b.size() * 4); cudaMemcpy(db[k], b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    p.W[k] = dW[k]; p.b[k] = db[k];
  }
  for (int it = 0; it < 3; ++it) {
    int st = molann_b200_forward(&p, dx, L, dy, nullptr, 0, nullptr);
    if (st) { printf("forward failed: %d\n", st); return 1; }
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  molann_b200_forward(&p, dx, L, dy, nullptr, 0, nullptr);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("# L=%d  %.3f ms  %.3f G frames/s  (%.0f cycles/tile/SM at 1.965 GHz)\n", L, ms, L / ms * 1e-6,
         ms * 1e-3 * 1.965e9 / (L / 128 / 148));
  static long long tr[8 * 64 * 8];
  cudaMemcpyFromSymbol(tr, molann::g_ws_trace, sizeof(tr));
  {
    const long long* k = &tr[7 * 64 * 8];
    printf("# CTA 0 lifetime: %lld SM cycles in %lld ns  ->  SM clock %.0f MHz while the kernel runs\n", k[2] - k[0], k[3] - k[1],
           1e3 * (double)(k[2] - k[0]) / (double)(k[3] - k[1]));
  }
  {
    static long long life[256 * 2];
    cudaMemcpyFromSymbol(life, molann::g_ws_life, sizeof(life));
    long long smin = life[0];
    for (int b = 0; b < 148; ++b) if (life[2 * b] < smin) smin = life[2 * b];
    long long smax = 0, emax = 0, emin = 1LL << 60; double esum = 0;
    for (int b = 0; b < 148; ++b) {
      const long long s0 = life[2 * b] - smin, e = life[2 * b + 1] - smin;
      if (s0 > smax) smax = s0; if (e > emax) emax = e; if (e < emin) emin = e; esum += e;
    }
    printf("# forward CTA lifetimes (ns from first CTA start): latest start %lld; end min %lld mean %.0f max %lld\n", smax, emin,
           esum / 148, emax);
    printf("# slowest / fastest CTAs:");
    for (int b = 0; b < 148; ++b) { const long long e = life[2 * b + 1] - smin; if (e > emax - (emax - emin) / 8 || e < emin + (emax - emin) / 8) printf(" %d:%lld", b, e); }
    printf("\n");
  }
  long long t0 = tr[(0 * 64 + 0) * 8 + 0];
  const char* names[7] = {"G0", "G1", "E1", "E2", "MMA1", "MMA2", "PROD"};
  for (int i = 0; i < 40; ++i) {
    printf("tile %2d:", i);
    for (int r = 0; r < 7; ++r) {
      printf("  %s", names[r]);
      for (int e = 0; e < 5; ++e) {
        long long v = tr[(r * 64 + i) * 8 + e];
        if (v) printf(" %lld", v - t0); else printf(" -");
      }
    }
    printf("\n");
  }
  // ---- fixed cost per launch: time vs tiles per SM ----
  {
    float *dgy2, *dgx2;
    cudaMalloc(&dgy2, (size_t)L * 2 * 4); cudaMalloc(&dgx2, x.size() * 4);
    cudaMemset(dgy2, 0, (size_t)L * 2 * 4);
    for (int k : {1, 2, 4, 8, 16, 40}) {
      const long long Lk = 148LL * 128 * k;
      float best_f = 1e9f, best_v = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        molann_b200_forward(&p, dx, Lk, dy, nullptr, 0, nullptr);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best_f) best_f = ms;
        cudaEventRecord(e0);
        molann_b200_value_and_grad(&p, dx, dgy2, Lk, dy, dgx2, nullptr, 0, nullptr);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best_v) best_v = ms;
      }
      printf("# tiles/SM %2d: forward %.1f us   value_and_grad %.1f us\n", k, best_f * 1e3, best_v * 1e3);
    }
  }
  // ---- value-and-gradient kernel (single-role, 2 tiles per CTA): phase stamps of warpgroup 0 of CTA 0 ----
  {
    float *dgy, *dgx;
    cudaMalloc(&dgy, (size_t)L * 2 * 4); cudaMalloc(&dgx, x.size() * 4);
    cudaMemset(dgy, 0, (size_t)L * 2 * 4);
    for (int it = 0; it < 3; ++it) molann_b200_value_and_grad(&p, dx, dgy, L, dy, dgx, nullptr, 0, nullptr);
    cudaEventRecord(e0);
    int st = molann_b200_value_and_grad(&p, dx, dgy, L, dy, dgx, nullptr, 0, nullptr);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("# value_and_grad status %d  %.3f ms  %.3f G frames/s\n", st, ms, L / ms * 1e-6);
    {
      static long long life[256 * 4];
      cudaMemcpyFromSymbol(life, molann::g_vg_life, sizeof(life));
      long long t0 = life[0], tmax = 0, smin = life[0];
      for (int b = 0; b < 148; ++b) { if (life[b * 4] < smin) smin = life[b * 4]; }
      double sum0 = 0, sum1 = 0; long long mx0 = 0, mx1 = 0, mn0 = 1LL << 60, smax = 0;
      for (int b = 0; b < 148; ++b) {
        const long long s0 = life[b * 4] - smin, e0 = life[b * 4 + 1] - smin, e1 = life[b * 4 + 2] - smin;
        if (s0 > smax) smax = s0;
        sum0 += e0; sum1 += e1;
        if (e0 > mx0) mx0 = e0; if (e1 > mx1) mx1 = e1; if (e0 < mn0) mn0 = e0;
      }
      printf("# vg CTA lifetimes (ns from first CTA start): latest start %lld; WG0 end mean %.0f max %lld min %lld; WG1 end mean %.0f max %lld\n",
             smax, sum0 / 148, mx0, mn0, sum1 / 148, mx1);
      (void)t0; (void)tmax;
    }
    static long long vt[64 * 16];
    cudaMemcpyFromSymbol(vt, molann::g_vg_trace, sizeof(vt));
    const char* ev[13] = {"start", "x", "kabsch", "feat", "mma1", "E1", "gyW", "mma2", "E2", "bwdMMA+E3", "lock", "zero", "featbwd"};
    for (int i = 4; i < 10; ++i) {
      printf("vg tile %2d:", i);
      for (int e = 1; e < 13; ++e) printf(" %s +%lld", ev[e], vt[i * 16 + e] - vt[i * 16 + e - 1]);
      printf("  | total %lld\n", vt[(i + 1) * 16] - vt[i * 16]);
    }
  }
  return 0;
}
