// fw_trace.cu -- DEVELOPMENT TOOL: role timelines (clock64, CTA 0) of the fused wide forward kernel on a C3-shaped
// problem (2000 atoms, every 10th aligned + position feature, 100 dihedrals, MLP [800,256,128,2]).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMOLANN_WS_TRACE -o fw_trace fw_trace.cu
#include "../../molann_b200/csrc/molann_b200.cu"

#include <cstdio>
#include <vector>

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2000;
  const int na = n / 10, nd = n / 20;
  const int reps = argc > 2 ? atoi(argv[2]) : 2;        // tiles per CTA
  const int L0 = 148 * 128, L = L0 * reps;             // host frames (one tile per CTA), replicated on the device
  std::vector<float> x((size_t)L0 * n * 3), ref(na * 3), base(n * 3);
  unsigned s = 777u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  for (int j = 0; j < n; ++j) { base[3 * j] = 1.5f * j * 0.3f + rnd(); base[3 * j + 1] = 10.f * rnd(); base[3 * j + 2] = 10.f * rnd(); }
  for (size_t f = 0; f < (size_t)L0; ++f)
    for (int j = 0; j < n * 3; ++j) x[f * n * 3 + j] = base[j] + 0.2f * rnd();
  std::vector<int> aidx(na), ent;
  float c[3] = {0, 0, 0};
  for (int k = 0; k < na; ++k) { aidx[k] = 10 * k; for (int d = 0; d < 3; ++d) c[d] += base[30 * k + d] / na; }
  for (int k = 0; k < na; ++k) for (int d = 0; d < 3; ++d) ref[3 * k + d] = base[30 * k + d] - c[d];
  int col = 0;
  for (int k = 0; k < na; ++k) { int e[6] = {3, 10 * k, 0, 0, 0, col}; ent.insert(ent.end(), e, e + 6); col += 3; }
  for (int k = 0; k < nd; ++k) { int e[6] = {2, 20 * k, 20 * k + 1, 20 * k + 2, 20 * k + 3, col}; ent.insert(ent.end(), e, e + 6); col += 2; }
  const int dims[4] = {col, 256, 128, 2};
  float *dx, *dy, *dref, *dW[3], *db[3]; int *daidx, *dent;
  cudaMalloc(&dx, x.size() * 4 * reps);
  for (int r = 0; r < reps; ++r) cudaMemcpy(dx + (size_t)r * x.size(), x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dy, (size_t)L * 2 * 4);
  cudaMalloc(&dref, ref.size() * 4); cudaMemcpy(dref, ref.data(), ref.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&daidx, na * 4); cudaMemcpy(daidx, aidx.data(), na * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dent, ent.size() * 4); cudaMemcpy(dent, ent.data(), ent.size() * 4, cudaMemcpyHostToDevice);
  MolannPlan p; std::memset(&p, 0, sizeof(p));
  p.n_inp = n; p.n_align = na; p.align_idx = daidx; p.ref_x = dref; p.n_entries = (int)ent.size() / 6; p.entries = dent;
  p.d_feat = col; p.use_angle_value = 0; p.n_layers = 3; p.act_id = 0;
  for (int k = 0; k < 3; ++k) {
    std::vector<float> w((size_t)dims[k] * dims[k + 1]), b(dims[k + 1]);
    for (auto& v : w) v = 2.f * rnd() / sqrtf((float)dims[k]);
    for (auto& v : b) v = 0.1f * rnd();
    cudaMalloc(&dW[k], w.size() * 4); cudaMemcpy(dW[k], w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&db[k], b.size() * 4); cudaMemcpy(db[k], b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    p.W[k] = dW[k]; p.b[k] = db[k]; p.dims[k] = dims[k];
  }
  p.dims[3] = dims[3];
  void* prep_buf; size_t pb = molann_b200_prepared_bytes(&p);
  cudaMalloc(&prep_buf, pb);
  MolannPrepared* h = nullptr;
  int st = molann_b200_prepare(&p, prep_buf, pb, nullptr, &h);
  printf("# prepare status %d (%zu bytes)\n", st, pb);
  size_t wsb = molann_b200_prepared_workspace_bytes(h, L);
  void* ws; cudaMalloc(&ws, wsb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    st = molann_b200_forward_prepared(h, &p, dx, L, dy, ws, wsb, nullptr);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("# forward_prepared status %d  %.3f ms for %d frames (%.1f M frames/s)  [%s]\n", st, ms, L, L / ms * 1e-3,
           cudaGetErrorString(cudaGetLastError()));
  }
#ifndef MOLANN_WS_TRACE
  return 0;                                   // timing-only build (ncu target)
#else
  static long long tr[4 * 256 * 8], tt[4 * 32 * 8];
  cudaMemcpyFromSymbol(tr, molann::g_fw_trace, sizeof(tr));
  cudaMemcpyFromSymbol(tt, molann::g_fw_tiles, sizeof(tt));
  {
    auto Q = [&](int role, int i, int ev) { return tt[(role * 32 + i) * 8 + ev]; };
    const long long q0 = Q(0, 0, 0);
    printf("per tile (CTA 0), cycles; waits are per-tile deltas of cumulative counters\n");
    for (int i = 0; i + 1 < reps && i < 31; ++i) {
      printf("tile %2d geo   @%9lld period %7lld | wait slot %7lld frame %7lld group-barrier %7lld\n", i, Q(0, i, 0) - q0,
             Q(0, i + 1, 0) - Q(0, i, 0), Q(0, i + 1, 1) - Q(0, i, 1), Q(0, i + 1, 2) - Q(0, i, 2), Q(0, i + 1, 3) - Q(0, i, 3));
      printf("        conv  @%9lld period %7lld | s_full %7lld prefetch+rotation %7lld turn %7lld chunks %7lld (of which wait empty %7lld)\n",
             Q(1, i, 0) - q0, Q(1, i + 1, 0) - Q(1, i, 0), Q(1, i, 1) - Q(1, i, 0), Q(1, i, 2) - Q(1, i, 1),
             Q(1, i, 3) - Q(1, i, 2), Q(1, i, 4) - Q(1, i, 3), Q(1, i, 5));
      printf("        mma   @%9lld period %7lld | layer-1 %7lld | wait a_full %7lld b_full %7lld d_free %7lld\n", Q(3, i, 0) - q0,
             Q(3, i + 1, 0) - Q(3, i, 0), Q(3, i, 4) - Q(3, i, 0), Q(3, i + 1, 1) - Q(3, i, 1), Q(3, i + 1, 2) - Q(3, i, 2),
             Q(3, i + 1, 3) - Q(3, i, 3));
      printf("        epi   @%9lld period %7lld | layer-1 sums %7lld h1 chunks + wait l2_full %7lld rest %7lld\n", Q(2, i, 0) - q0,
             Q(2, i + 1, 0) - Q(2, i, 0), Q(2, i, 1) - Q(2, i, 0), Q(2, i, 3) - Q(2, i, 1), Q(2, i, 4) - Q(2, i, 3));
    }
  }
  auto T = [&](int role, int i, int ev) { return tr[(role * 256 + i) * 8 + ev]; };
  const long long t00 = T(0, 0, 0);
  printf("geometry warp 0 (frames it owns; cycles): wait slot | wait frame | moment loop | staging free + reduce | positions | "
         "invariants | fence | store + arrive | total\n");
  for (int i = 0; i < 232; i += (i == 23 ? 177 : 1))
    printf("g %3d @%8lld: slot+%lld frame+%lld mom+%lld red+%lld pos+%lld inv+%lld fence+%lld out+%lld | %lld\n", i,
           T(0, i, 0) - t00, T(0, i, 1) - T(0, i, 0), T(0, i, 2) - T(0, i, 1), T(0, i, 3) - T(0, i, 2),
           T(0, i, 4) - T(0, i, 3), T(0, i, 5) - T(0, i, 4), T(0, i, 6) - T(0, i, 5), T(0, i, 7) - T(0, i, 6),
           T(0, i + 1, 0) - T(0, i, 7), T(0, i + 1, 0) - T(0, i, 0));
  printf("MMA warp, tile 0 layer-1 chunks: wait d_free | wait a_full | wait b_full | issue+commit | period\n");
  for (int i = 0; i < 20; ++i)
    printf("m %2d @%8lld: dfree+%lld a+%lld b+%lld issue+%lld | %lld\n", i, T(3, i, 0) - t00, T(3, i, 1) - T(3, i, 0),
           T(3, i, 2) - T(3, i, 1), T(3, i, 3) - T(3, i, 2), T(3, i, 4) - T(3, i, 3), T(3, i + 1, 0) - T(3, i, 0));
  printf("converter warp 0, tile 0 chunks: (load+rotate) | wait empty | split+store | fence+arrive\n");
  for (int i = 0; i < 20; ++i)
    printf("c %2d @%8lld: prep+%lld empty+%lld store+%lld arrive+%lld\n", i, T(1, 16 + i, 0) - t00,
           i ? T(1, 16 + i, 0) - T(1, 15 + i, 3) : 0LL, T(1, 16 + i, 1) - T(1, 16 + i, 0), T(1, 16 + i, 2) - T(1, 16 + i, 1),
           T(1, 16 + i, 3) - T(1, 16 + i, 2));
  printf("epilogue warp 4, tile 0 segments: wait d_full | drain\n");
  for (int i = 0; i < 12; ++i)
    printf("e %2d @%8lld: wait+%lld drain+%lld\n", i, T(2, 16 + i, 0) - t00, T(2, 16 + i, 1) - T(2, 16 + i, 0),
           T(2, 16 + i, 2) - T(2, 16 + i, 1));
  for (int i = 0; i < 2; ++i)
    printf("converter tile %d @%8lld: turn+%lld s_full+%lld rotation+%lld chunks+%lld\n", i, T(1, i, 0) - t00,
           T(1, i, 1) - T(1, i, 0), T(1, i, 2) - T(1, i, 1), T(1, i, 3) - T(1, i, 2), T(1, i, 4) - T(1, i, 3));
  for (int i = 0; i < 2; ++i)
    printf("epilogue  tile %d @%8lld: layer1 sums+%lld h1 chunks+%lld l2_full+%lld rest+%lld\n", i, T(2, i, 0) - t00,
           T(2, i, 1) - T(2, i, 0), T(2, i, 2) - T(2, i, 1), T(2, i, 3) - T(2, i, 2), T(2, i, 4) - T(2, i, 3));
  return 0;
#endif
}
