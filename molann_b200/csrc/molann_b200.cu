// molann_b200.cu -- C ABI (include/molann_b200.h) and kernel dispatch for the molann hot path.
//
// Two kernel families:
//   * fused small-system kernels (fused_small.cuh): one kernel, frame tile + MLP resident in smem;
//   * general path (general.cuh): warp-per-frame geometry + layered FFMA GEMMs over a caller
//     workspace, processed in frame chunks so scratch stays O(chunk).
// No CPU fallback exists: without a CUDA device every entry point returns MOLANN_ERR_CUDA.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "fused_small.cuh"
#include "fused_tc.cuh"
#include "fused_ws.cuh"
#include "fused_wide.cuh"
#include "gemm_tc.cuh"
#include "staged_block.cuh"
#include "small_tile.cuh"
#include "general.cuh"
#include "fused_train.cuh"

using namespace molann;

namespace {

std::atomic<long long> g_launches{0};
thread_local int t_last_cuda_error = 0;

struct DeviceInfo {
  int sm_count = 0;
  int max_smem_optin = 0;
  int smem_per_sm = 0;
  bool ok = false;
};

// The attributes of a device never change: queried once per device and cached (an MD plugin calls with a handful
// of frames, where three attribute queries per call are a measurable part of the host-side cost).  The CURRENT
// device is still asked for on every call, so switching devices between calls stays correct.
DeviceInfo device_info() {
  constexpr int kMaxDev = 64;
  static DeviceInfo cache[kMaxDev];
  static std::atomic<int> ready[kMaxDev];
  DeviceInfo d;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return d;
  const bool cacheable = dev >= 0 && dev < kMaxDev;
  if (cacheable && ready[dev].load(std::memory_order_acquire)) return cache[dev];
  if (cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return d;
  if (cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return d;
  if (cudaDeviceGetAttribute(&d.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess) return d;
  d.ok = true;
  if (cacheable) {                       // racing first calls write identical values
    cache[dev] = d;
    ready[dev].store(1, std::memory_order_release);
  }
  return d;
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

int check_cuda(cudaError_t e) {
  if (e == cudaSuccess) return MOLANN_OK;
  t_last_cuda_error = (int)e;
  return MOLANN_ERR_CUDA;
}

int post_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError());
}

DevPlan to_dev(const MolannPlan* p) {
  DevPlan d;
  std::memset(&d, 0, sizeof(d));
  d.n_inp = p->n_inp; d.n_align = p->n_align; d.n_entries = p->n_entries; d.d_feat = p->d_feat;
  d.use_angle = p->use_angle_value; d.n_layers = p->n_layers; d.act = p->act_id;
  d.align_idx = p->align_idx; d.ref_x = p->ref_x; d.entries = p->entries;
  for (int k = 0; k <= MOLANN_MAX_LAYERS; ++k) d.dims[k] = p->dims[k];
  for (int k = 0; k < MOLANN_MAX_LAYERS; ++k) { d.W[k] = p->W[k]; d.b[k] = p->b[k]; }
  return d;
}

int validate_geometry(const MolannPlan* p) {
  if (!p) return MOLANN_ERR_NULL;
  if (p->n_inp <= 0 || p->n_align < 0 || p->n_align > p->n_inp * 64) return MOLANN_ERR_PLAN;
  if (p->n_align > 0 && (!p->align_idx || !p->ref_x)) return MOLANN_ERR_NULL;
  return MOLANN_OK;
}

int validate_features(const MolannPlan* p) {
  int s = validate_geometry(p);
  if (s) return s;
  if (p->n_entries <= 0 || p->d_feat <= 0) return MOLANN_ERR_PLAN;
  if (!p->entries) return MOLANN_ERR_NULL;
  if (p->use_angle_value != 0 && p->use_angle_value != 1) return MOLANN_ERR_PLAN;
  return MOLANN_OK;
}

int validate_full(const MolannPlan* p) {
  int s = validate_features(p);
  if (s) return s;
  if (p->n_layers < 0 || p->n_layers > MOLANN_MAX_LAYERS) return MOLANN_ERR_PLAN;
  if (p->n_layers > 0) {
    if (p->act_id < 0 || p->act_id > MOLANN_ACT_IDENTITY) return MOLANN_ERR_PLAN;
    if (p->dims[0] != p->d_feat) return MOLANN_ERR_PLAN;
    for (int k = 0; k <= p->n_layers; ++k)
      if (p->dims[k] <= 0) return MOLANN_ERR_PLAN;
    for (int k = 0; k < p->n_layers; ++k)
      if (!p->W[k] || !p->b[k]) return MOLANN_ERR_NULL;
  }
  return MOLANN_OK;
}

bool misaligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) != 0; }

// ---------------------------------------------------------------------------------------------
// fused small-system path
// ---------------------------------------------------------------------------------------------
int pick_tm(int F, int NT, int n_out) {
  const int nog = (n_out + 7) / 8;
  const int cands[4] = {8, 4, 2, 1};
  for (int c = 0; c < 4; ++c)
    if ((F / cands[c]) * nog >= NT) return cands[c];
  return 1;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct Carver {
  int off = 0;
  int take(int bytes, int align = 16) {
    off = (off + align - 1) / align * align;
    const int r = off;
    off += bytes;
    return r;
  }
};

SmallLayout small_layout(const MolannPlan* p, int F, int NT, bool backward, bool alias) {
  SmallLayout lay;
  std::memset(&lay, 0, sizeof(lay));
  const int n3 = 3 * p->n_inp;
  const int nl = p->n_layers;
  Carver c;
  lay.mbar_off = c.take(16, 16);
  const int xs_bytes = F * n3 * 4;
  if (!backward) {
    int rows0 = round_up(p->d_feat, 8), rows1 = 8;
    for (int k = 1; k < nl; ++k) {       // h_k lands in buffer k & 1
      const int r = round_up(p->dims[k], 8);
      if (k & 1) rows1 = r > rows1 ? r : rows1; else rows0 = r > rows0 ? r : rows0;
    }
    const int b1_bytes = rows1 * F * 4;
    if (alias) {                         // activation buffer 1 overlays the coordinate tile
      lay.alias_xs = 1;
      lay.xs_off = c.take(xs_bytes > b1_bytes ? xs_bytes : b1_bytes, 128);
      lay.buf_off[1] = lay.xs_off;
    } else {
      lay.xs_off = c.take(xs_bytes, 128);
      lay.buf_off[1] = c.take(b1_bytes, 128);
    }
    lay.buf_off[0] = c.take(rows0 * F * 4, 128);
    for (int k = 0; k < nl; ++k) {
      lay.ldwt[k] = round_up(p->dims[k + 1], 8);
      lay.wt_off[k] = c.take(p->dims[k] * lay.ldwt[k] * 4, 16);
      lay.b_off[k] = c.take(lay.ldwt[k] * 4, 16);
      lay.tm_fwd[k] = pick_tm(F, NT, p->dims[k + 1]);
    }
  } else {
    lay.xs_off = c.take(xs_bytes, 128);
    lay.gxs_off = c.take(xs_bytes, 128);
    int rows[MOLANN_MAX_LAYERS];
    for (int s = 0; s < MOLANN_MAX_LAYERS; ++s) rows[s] = 0;
    rows[0] = round_up(p->d_feat, 8);
    for (int k = 1; k < nl; ++k) {
      const int s = act_slot(k), r = round_up(p->dims[k], 8);
      rows[s] = r > rows[s] ? r : rows[s];
    }
    for (int s = 0; s < MOLANN_MAX_LAYERS; ++s)
      if (rows[s] > 0) lay.buf_off[s] = c.take(rows[s] * F * 4, 128);
    lay.gys_off = c.take(round_up(p->dims[nl], 8) * F * 4, 128);
    for (int k = 0; k < nl; ++k) {
      lay.ldw[k] = round_up(p->dims[k], 8);
      lay.w_off[k] = c.take(round_up(p->dims[k + 1], 8) * lay.ldw[k] * 4, 16);
      lay.tm_bwd[k] = pick_tm(F, NT, p->dims[k]);
      if (k < nl - 1) {
        lay.ldwt[k] = round_up(p->dims[k + 1], 8);
        lay.wt_off[k] = c.take(p->dims[k] * lay.ldwt[k] * 4, 16);
        lay.b_off[k] = c.take(lay.ldwt[k] * 4, 16);
        lay.tm_fwd[k] = pick_tm(F, NT, p->dims[k + 1]);
      }
    }
  }
  lay.aidx_off = c.take((p->n_align > 0 ? p->n_align : 1) * 4, 16);
  lay.ref_off = c.take((p->n_align > 0 ? 3 * p->n_align : 1) * 4, 16);
  lay.ent_off = c.take(p->n_entries * ENTRY_INTS * 4, 16);
  lay.total_bytes = round_up(c.off, 128);
  return lay;
}

struct SmallChoice {
  bool ok = false;
  int F = 0, NT = 0;
  bool alias = false;
  SmallLayout lay;
};

// Tunables (env, read per call; defaults chosen from B200 measurements):
//   MOLANN_B200_PATH      = 0 general | 1 fused small | -1 auto (default)
//   MOLANN_B200_FWD_F/NT  frames per tile / threads per CTA of the fused forward kernel
//   MOLANN_B200_BWD_F/NT  same for the fused backward kernel
//   MOLANN_B200_FWD_ALIAS overlay activation buffer 1 on the coordinate tile (forward)
SmallChoice choose_small(const MolannPlan* p, bool backward, const DeviceInfo& dev) {
  SmallChoice ch;
  if (p->n_layers < 1) return ch;
  const int forced = env_int("MOLANN_B200_PATH", -1);
  if (forced == 0) return ch;
  if ((long long)p->n_entries * ENTRY_INTS * 4 > 64 * 1024) return ch;
  int F = backward ? env_int("MOLANN_B200_BWD_F", 64) : env_int("MOLANN_B200_FWD_F", 128);
  int NT = backward ? env_int("MOLANN_B200_BWD_NT", 128) : env_int("MOLANN_B200_FWD_NT", 128);
  if (F != 64 && F != 128) F = backward ? 64 : 128;
  if (NT != 128 && NT != 256) NT = 128;
  const bool alias = !backward && env_int("MOLANN_B200_FWD_ALIAS", 1) != 0;
  // try the preferred tile first, then the smaller one
  const int tries[2] = {F, 64};
  for (int t = 0; t < 2; ++t) {
    const int Ft = tries[t];
    if ((long long)Ft * 3 * p->n_inp * 4 > 200 * 1024) continue;
    SmallLayout lay = small_layout(p, Ft, NT, backward, alias);
    if (lay.total_bytes <= dev.max_smem_optin) {
      ch.ok = true; ch.F = Ft; ch.NT = NT; ch.alias = alias; ch.lay = lay;
      return ch;
    }
  }
  return ch;
}

template <int F, int NT>
int launch_small_forward(const DevPlan& dp, const SmallLayout& lay, const float* x, float* y, long long L,
                         const DeviceInfo& dev, cudaStream_t st) {
  auto kern = fused_small_forward_kernel<F, NT>;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total_bytes));
  if (s) return s;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, lay.total_bytes);
  if (occ < 1) occ = 1;
  const long long ntiles = (L + F - 1) / F;
  long long grid = (long long)dev.sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((F * 3 * dp.n_inp * 4) % 16 == 0);
  kern<<<(unsigned)grid, NT, lay.total_bytes, st>>>(dp, lay, x, y, L, use_tma);
  return post_launch();
}

template <int F, int NT>
int launch_small_backward(const DevPlan& dp, const SmallLayout& lay, const float* x, const float* gy, float* gx,
                          long long L, const DeviceInfo& dev, cudaStream_t st) {
  auto kern = fused_small_backward_kernel<F, NT>;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total_bytes));
  if (s) return s;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, lay.total_bytes);
  if (occ < 1) occ = 1;
  const long long ntiles = (L + F - 1) / F;
  long long grid = (long long)dev.sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15u) == 0) &&
                      ((F * 3 * dp.n_inp * 4) % 16 == 0);
  kern<<<(unsigned)grid, NT, lay.total_bytes, st>>>(dp, lay, x, gy, gx, L, use_tma);
  return post_launch();
}

// ---------------------------------------------------------------------------------------------
// fused tensor-core path (tcgen05 3xTF32 MLP), forward
// ---------------------------------------------------------------------------------------------
struct TcChoice {
  bool ok = false;
  TcLayout lay;
};

// MOLANN_B200_TC = 0 disables the tensor-core kernels (A/B testing against the FFMA kernels).
TcChoice choose_tc(const MolannPlan* p, bool backward, const DeviceInfo& dev) {
  TcChoice ch;
  std::memset(&ch.lay, 0, sizeof(ch.lay));
  if (backward) return ch;
  if (env_int("MOLANN_B200_TC", 1) == 0 || env_int("MOLANN_B200_PATH", -1) == 0) return ch;
  const int nl = p->n_layers;
  if (nl < 2) return ch;
  for (int k = 0; k < nl; ++k)
    if (p->dims[k] > TC_MAXW) return ch;
  if (p->dims[nl] > 8) return ch;
  if ((long long)p->n_entries * ENTRY_INTS * 4 > 32 * 1024) return ch;
  if ((long long)TC_F * 3 * p->n_inp * 4 > 96 * 1024) return ch;
  TcLayout& lay = ch.lay;
  Carver c;
  lay.mbar_off = c.take(16, 16);
  lay.tptr_off = c.take(16, 16);
  lay.xs_off = c.take(TC_F * 3 * p->n_inp * 4, 128);
  for (int k = 0; k < nl - 1; ++k) {
    lay.kp[k] = round_up(p->dims[k], 16);
    lay.np[k] = round_up(p->dims[k + 1], 16);
    lay.bhi_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 1024);
    lay.blo_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 1024);
    lay.bias_off[k] = c.take(lay.np[k] * 4, 16);
  }
  lay.wlast_off = c.take(p->dims[nl] * TC_MAXW * 4, 16);
  lay.blast_off = c.take(p->dims[nl] * 4, 16);
  lay.aidx_off = c.take((p->n_align > 0 ? p->n_align : 1) * 4, 16);
  lay.ref_off = c.take((p->n_align > 0 ? 3 * p->n_align : 1) * 4, 16);
  lay.ent_off = c.take(p->n_entries * ENTRY_INTS * 4, 16);
  lay.total_bytes = round_up(c.off, 128);
  // exactly two CTAs (2 x 256 TMEM columns) may share an SM: keep the footprint above a third of it
  const int min_bytes = dev.max_smem_optin / 3 + 1024;
  if (lay.total_bytes < min_bytes) lay.total_bytes = round_up(min_bytes, 128);
  if (lay.total_bytes > dev.max_smem_optin) return ch;
  ch.ok = true;
  return ch;
}

int launch_tc_forward(const DevPlan& dp, const TcLayout& lay, const float* x, float* y, long long L,
                      const DeviceInfo& dev, cudaStream_t st) {
  auto kern = fused_tc_forward_kernel;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total_bytes));
  if (s) return s;
  // ask for the full shared-memory carveout so that two CTAs (2 x 256 TMEM columns) share an SM
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = dev.smem_per_sm / (lay.total_bytes + 1024);
  if (occ < 1) occ = 1;
  if (occ > 2) occ = 2;
  const long long ntiles = (L + TC_F - 1) / TC_F;
  long long grid = (long long)dev.sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((TC_F * 3 * dp.n_inp * 4) % 16 == 0);
  kern<<<(unsigned)grid, TC_F, lay.total_bytes, st>>>(dp, lay, x, y, L, use_tma);
  return post_launch();
}

// ---------------------------------------------------------------------------------------------
// warp-specialised fused forward (fused_ws.cuh): the default for the small-system class
// ---------------------------------------------------------------------------------------------
struct WsChoice {
  bool ok = false;
  WsLayout wl;
};

// MOLANN_B200_WS = 0 falls back to the single-role tensor-core kernel (A/B testing).
WsChoice choose_ws(const MolannPlan* p, const float* x, const DeviceInfo& dev) {
  WsChoice ch;
  std::memset(&ch.wl, 0, sizeof(ch.wl));
  if (env_int("MOLANN_B200_WS", 1) == 0 || env_int("MOLANN_B200_TC", 1) == 0 || env_int("MOLANN_B200_PATH", -1) == 0)
    return ch;
  const int nl = p->n_layers;
  if (nl < 2 || nl > 3) return ch;
  for (int k = 0; k < nl; ++k)
    if (p->dims[k] > TC_MAXW) return ch;
  if (p->dims[nl] > 8) return ch;
  if ((long long)p->n_entries * ENTRY_INTS * 4 > 32 * 1024) return ch;
  (void)x;                                            // any 4-byte aligned x: the kernel copies from the 16-byte
                                                      // boundary below each tile
  const long long tile_bytes = (long long)WS_F * 3 * p->n_inp * 4;
  WsLayout& wl = ch.wl;
  TcLayout& lay = wl.base;
  Carver c;
  wl.mbar_off = c.take(8 * 32, 16);
  wl.tptr_off = c.take(16, 16);
  for (int k = 0; k < nl - 1; ++k) {
    lay.kp[k] = round_up(p->dims[k], 16);
    lay.np[k] = round_up(p->dims[k + 1], 16);
    lay.bhi_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 128);
    lay.blo_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 128);
    lay.bias_off[k] = 0;                                  // the bias goes through the MMA (bbh / bbl below)
    wl.bbh_off[k] = c.take(2 * lay.np[k] * 16, 128);
    wl.bbl_off[k] = c.take(2 * lay.np[k] * 16, 128);
  }
  lay.wlast_off = c.take(p->dims[nl] * TC_MAXW * 4, 16);
  lay.blast_off = c.take(p->dims[nl] * 4, 16);
  wl.aoff_off = c.take((p->n_align > 0 ? p->n_align : 1) * 4, 16);
  wl.ref4_off = c.take((p->n_align > 0 ? p->n_align : 1) * 16, 16);
  lay.ent_off = c.take(p->n_entries * ENTRY_INTS * 4, 16);
  wl.ones_off = c.take(2 * WS_F * 16, 128);
  for (int b = 0; b < 2; ++b) wl.a1s_off[b] = c.take(2 * round_up(p->dims[0], 16) * WS_F * 4, 128);
  // coordinate-tile ring: 3 deep when it fits, else 2 (the kernel then exposes a little more load latency, which is
  // still far better than falling back to the single-role kernel)
  wl.n_xbuf = 0;
  for (int b = 0; b < WS_XBUF; ++b) {
    if (c.off + tile_bytes + 16 + 128 > dev.max_smem_optin) break;
    wl.xs_off[b] = c.take((int)tile_bytes + 16, 128);
    wl.n_xbuf = b + 1;
  }
  if (wl.n_xbuf < 2) return ch;
  wl.total_bytes = round_up(c.off, 128);
  if (wl.total_bytes > dev.max_smem_optin) return ch;
  // TMEM columns (A1 lives in shared memory): in order of value, double-buffered accumulators, second A2 buffer
  const int np0 = lay.np[0];
  const int kp1 = nl == 3 ? lay.kp[1] : 0, np1 = nl == 3 ? lay.np[1] : 0;
  const int fixed = 2 * kp1;
  const int dcols = np0 + np1;
  wl.n_dbuf = (fixed + 2 * dcols <= 512) ? 2 : 1;
  wl.n_a2buf = (nl == 3 && fixed + wl.n_dbuf * dcols + 2 * kp1 <= 512) ? 2 : 1;
  if (fixed + dcols > 512) return ch;
  int col = 0;
  if (nl == 3) {
    wl.col_a2[0] = col; col += 2 * kp1;
    wl.col_a2[1] = wl.col_a2[0];
    if (wl.n_a2buf == 2) { wl.col_a2[1] = col; col += 2 * kp1; }
    for (int b = 0; b < 2; ++b) {
      if (b < wl.n_dbuf) { wl.col_d1[b] = col; col += np0; } else wl.col_d1[b] = wl.col_d1[0];
    }
    for (int b = 0; b < 2; ++b) {
      if (b < wl.n_dbuf) { wl.col_d2[b] = col; col += np1; } else wl.col_d2[b] = wl.col_d2[0];
    }
  } else {
    for (int b = 0; b < 2; ++b) {
      if (b < wl.n_dbuf) { wl.col_d2[b] = col; col += np0; } else wl.col_d2[b] = wl.col_d2[0];
      wl.col_d1[b] = wl.col_d2[b];
    }
  }
  if (col > 512) return ch;
  wl.tmem_cols = 512;
  ch.ok = true;
  return ch;
}

template <int ACT>
int launch_ws_forward_act(const DevPlan& dp, const WsLayout& wl, const float* x, float* y, long long L,
                          const DeviceInfo& dev, cudaStream_t st) {
  auto kern = fused_ws_forward_kernel<ACT>;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, wl.total_bytes));
  if (s) return s;
  const long long ntiles = (L + WS_F - 1) / WS_F;
  long long grid = dev.sm_count;                       // persistent: one CTA per SM
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, WS_THREADS, wl.total_bytes, st>>>(dp, wl, x, y, L);
  return post_launch();
}

int launch_ws_forward(const DevPlan& dp, const WsLayout& wl, const float* x, float* y, long long L,
                      const DeviceInfo& dev, cudaStream_t st) {
  switch (dp.act) {
    case ACT_TANH: return launch_ws_forward_act<ACT_TANH>(dp, wl, x, y, L, dev, st);
    case ACT_RELU: return launch_ws_forward_act<ACT_RELU>(dp, wl, x, y, L, dev, st);
    case ACT_SIGMOID: return launch_ws_forward_act<ACT_SIGMOID>(dp, wl, x, y, L, dev, st);
    default: return launch_ws_forward_act<ACT_IDENTITY>(dp, wl, x, y, L, dev, st);
  }
}

// ---------------------------------------------------------------------------------------------
// fused tensor-core value-and-gradient path (forward recompute + d/dx in one kernel)
// ---------------------------------------------------------------------------------------------
struct TcVgChoice {
  bool ok = false;
  int tiles = 0;
  TcVgLayout vl;
};

TcVgChoice choose_tc_vg(const MolannPlan* p, const DeviceInfo& dev, int force_tiles = 0) {
  TcVgChoice ch;
  std::memset(&ch.vl, 0, sizeof(ch.vl));
  if (env_int("MOLANN_B200_TC", 1) == 0 || env_int("MOLANN_B200_PATH", -1) == 0) return ch;
  const int nl = p->n_layers;
  if (nl < 2 || nl > 3) return ch;                 // one or two tensor-core layers (h_1 parked in TMEM)
  for (int k = 0; k < nl; ++k)
    if (p->dims[k] > TC_MAXW) return ch;
  if (p->dims[nl] > 8) return ch;
  if ((long long)p->n_entries * ENTRY_INTS * 4 > 32 * 1024) return ch;
  const int xs_bytes = TC_F * 3 * p->n_inp * 4;
  if (xs_bytes > 64 * 1024) return ch;
  const int want_tiles = force_tiles > 0 ? force_tiles : env_int("MOLANN_B200_VG_TILES", 2);
  for (int tiles = (want_tiles == 1 ? 1 : 2); tiles >= 1; --tiles) {
    TcVgLayout vl;
    std::memset(&vl, 0, sizeof(vl));
    TcLayout& lay = vl.base;
    Carver c;
    vl.mbar_off = c.take(16 * 4, 16);
    vl.tptr_off = c.take(16, 16);
    for (int k = 0; k < nl - 1; ++k) {
      lay.kp[k] = round_up(p->dims[k], 16);
      lay.np[k] = round_up(p->dims[k + 1], 16);
      lay.bhi_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 1024);
      lay.blo_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 1024);
      lay.bias_off[k] = c.take(lay.np[k] * 4, 16);
    }
    lay.wlast_off = c.take(p->dims[nl] * TC_MAXW * 4, 16);
    lay.blast_off = c.take(p->dims[nl] * 4, 16);
    lay.aidx_off = c.take((p->n_align > 0 ? p->n_align : 1) * 4, 16);
    lay.ref_off = c.take((p->n_align > 0 ? 3 * p->n_align : 1) * 4, 16);
    lay.ent_off = c.take(p->n_entries * ENTRY_INTS * 4, 16);
    for (int k = 0; k < nl - 1; ++k) {             // W^T operands for the backward contractions
      lay.thi_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 1024);
      lay.tlo_off[k] = c.take(lay.kp[k] * lay.np[k] * 4, 1024);
    }
    vl.lock_off = c.take(16, 16);
    vl.gxs_off = c.take(xs_bytes, 128);            // ONE gradient tile shared by the warpgroups
    for (int t = 0; t < tiles; ++t) vl.xs_off[t] = c.take(xs_bytes, 128);
    vl.total_bytes = round_up(c.off, 128);
    if (tiles == 1) {        // 256 TMEM columns per CTA: at most two CTAs may share an SM
      const int min_bytes = dev.max_smem_optin / 3 + 1024;
      if (vl.total_bytes < min_bytes) vl.total_bytes = round_up(min_bytes, 128);
    } else {                 // 512 TMEM columns: exactly one CTA per SM
      const int min_bytes = dev.max_smem_optin / 2 + 1024;
      if (vl.total_bytes < min_bytes) vl.total_bytes = round_up(min_bytes, 128);
    }
    if (vl.total_bytes <= dev.max_smem_optin) {
      ch.ok = true; ch.tiles = tiles; ch.vl = vl;
      return ch;
    }
  }
  return ch;
}

template <int TILES, int ACT>
int launch_tc_vg(const DevPlan& dp, const TcVgLayout& vl, const float* x, const float* gy, float* y, float* gx,
                 long long L, const DeviceInfo& dev, cudaStream_t st) {
  auto kern = fused_tc_value_grad_kernel<TILES, ACT>;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, vl.total_bytes));
  if (s) return s;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = dev.smem_per_sm / (vl.total_bytes + 1024);
  const int occ_cap = (TILES == 1) ? 2 : 1;
  if (occ < 1) occ = 1;
  if (occ > occ_cap) occ = occ_cap;
  const long long ntiles = (L + TC_F - 1) / TC_F;
  long long grid = (long long)dev.sm_count * occ;
  const long long need = (ntiles + TILES - 1) / TILES;
  if (grid > need) grid = need;
  const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15u) == 0) &&
                      ((TC_F * 3 * dp.n_inp * 4) % 16 == 0);
  kern<<<(unsigned)grid, TILES * TC_F, vl.total_bytes, st>>>(dp, vl, x, gy, y, gx, L, use_tma);
  return post_launch();
}

int run_tc_vg(const TcVgChoice& ch, const DevPlan& dp, const float* x, const float* gy, float* y, float* gx,
              long long L, const DeviceInfo& dev, cudaStream_t st) {
#define VG_CASE(A)                                                              \
  case A:                                                                       \
    return ch.tiles == 2 ? launch_tc_vg<2, A>(dp, ch.vl, x, gy, y, gx, L, dev, st) \
                         : launch_tc_vg<1, A>(dp, ch.vl, x, gy, y, gx, L, dev, st);
  switch (dp.act) {
    VG_CASE(ACT_TANH)
    VG_CASE(ACT_RELU)
    VG_CASE(ACT_SIGMOID)
    default:
      return ch.tiles == 2 ? launch_tc_vg<2, ACT_IDENTITY>(dp, ch.vl, x, gy, y, gx, L, dev, st)
                           : launch_tc_vg<1, ACT_IDENTITY>(dp, ch.vl, x, gy, y, gx, L, dev, st);
  }
#undef VG_CASE
}


// ---------------------------------------------------------------------------------------------
// fused wide path (fused_wide.cuh): prepared plans
// ---------------------------------------------------------------------------------------------
}  // namespace

struct MolannPrepared {
  uint32_t magic;
  // sizes the plan must still have when the handle is used
  int n_inp, n_align, n_entries, d_feat, n_layers, act, dims[MOLANN_MAX_LAYERS + 1];
  // kernel-order program
  int n_pos, n_inv_ent, n_inv, n_units, n_hidden, pos_is_align;
  int nkc1, n1, n1p, nkc2, n2, n2p, nlast, nlastp, kout, kpad;
  int ku;                    // K-chunk of the fused wide kernel in 16-byte units (fused_wide.cuh): 4, or 2 for very large frames
  // device pointers into the caller's buffer
  int* pos_atom;
  int* inv_ent;
  int* colmap;
  float* w1p;
  float* w2p;
  float* b1s;
  float* b2s;
  float* w3;
  float* b3;
  // layered tensor-core GEMMs (gemm_tc.cuh: value-and-gradient of these plans, forward when the wide kernel does not
  // fit): operands packed once -- forward orientation W_k and the backward-to-input orientation W_k^T; NULL where the
  // layer is too narrow for that kernel
  float* gt_fwd[MOLANN_MAX_LAYERS];
  float* gt_bwd[MOLANN_MAX_LAYERS];
};

namespace {

constexpr uint32_t kPreparedMagic = 0x4d4c5750u;   // "MLWP"

// MOLANN_B200_WIDE = 0 never takes the fused wide kernel, = 1 takes it for every plan of a supported shape (tests),
// default: plans no small-system kernel serves.
bool wide_shape_ok(const MolannPlan* p) {
  if (validate_full(p) != MOLANN_OK) return false;
  const int nl = p->n_layers;
  if (nl != 2 && nl != 3) return false;
  if (p->dims[1] > FW_NMAX) return false;
  if (nl == 3 && p->dims[2] > FW_NMAX) return false;
  if (p->dims[nl] > 8) return false;
  if (12LL * p->n_inp + 64 > 100 * 1024) return false;          // two frames must fit next to the operand ring
  return true;
}

// K-chunk of the fused wide kernel: 16 (KU = 4) unless three frames of the ring and two full operand stages do not fit
// next to each other in shared memory (C5: 60 KB frames) -- then 8 (KU = 2), which halves the stages.
int wide_ku_for(const MolannPlan* p) {
  const int forced = env_int("MOLANN_B200_WIDE_KU", 0);          // tests: both chunk sizes over the same goldens
  if (forced == 2 || forced == 4) return forced;
  const int ring_slot = round_up(12 * p->n_inp + 32, 128);
  const int budget = 227 * 1024;
  return 2 * fw_stage_bytes(4) + 2 * fw_conv_chunk(4) + 8192 + 3 * ring_slot <= budget ? 4 : 2;
}

struct WideCounts {
  int n_pos = 0, n_inv_ent = 0, n_inv = 0, n_units = 0;
};
int wide_units(int n_pos, int n_inv) { return n_pos + (n_inv > n_pos ? (n_inv - n_pos + 3) / 4 : 0); }

size_t wide_prepared_bytes_for(const MolannPlan* p, int n_pos_max, int n_inv_ent_max, int n_units_max) {
  const int nl = p->n_layers;
  const int n1p = round_up(p->dims[1], 16);
  const int FW_KU = wide_ku_for(p), FW_KC = fw_kc(FW_KU);
  const int nkc1 = (n_units_max + FW_KU - 1) / FW_KU;
  size_t b = 0;
  b += align256((size_t)n_pos_max * 4);
  b += align256((size_t)n_inv_ent_max * ENTRY_INTS * 4);
  b += align256((size_t)nkc1 * FW_KC * 4);                                  // colmap
  b += align256((size_t)nkc1 * 2 * FW_KC * n1p * 4);                        // w1p
  b += align256((size_t)n1p * 4);
  if (nl == 3) {
    const int n2p = round_up(p->dims[2], 16);
    b += align256((size_t)(n1p / FW_KC) * 2 * FW_KC * n2p * 4);             // w2p
    b += align256((size_t)n2p * 4);
    b += align256((size_t)p->dims[nl] * n2p * 4);
  } else {
    b += align256((size_t)p->dims[nl] * n1p * 4);
  }
  b += align256((size_t)p->dims[nl] * 4);
  for (int k = 0; k < nl; ++k) {
    if (p->dims[k] >= 32 && p->dims[k + 1] >= 32) {
      b += align256((size_t)gt_pack_floats(p->dims[k + 1], p->dims[k]) * 4);
      b += align256((size_t)gt_pack_floats(p->dims[k], p->dims[k + 1]) * 4);
    }
  }
  return b;
}

int launch_gemm_pack(const float* W, long long rs, long long cs, int N, int K, float* out, cudaStream_t st) {
  const long long total = (long long)((N + GT_NMAX - 1) / GT_NMAX) * GT_NMAX * round_up(K, GT_KC);
  unsigned pb = (unsigned)((total + 255) / 256);
  if (pb > 4096u) pb = 4096u;
  gemm_tc_pack_kernel<<<pb, 256, 0, st>>>(W, rs, cs, N, K, out);
  return post_launch();
}

int wide_pack_weights(const MolannPrepared* h, const MolannPlan* p, cudaStream_t st) {
  const float scale = ws_scale_for_act(p->act_id);
  {
    const long long total = (long long)h->n1p * h->kpad;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 2048u) blocks = 2048u;
    fw_pack_kernel<<<blocks, 256, 0, st>>>(p->W[0], p->d_feat, h->n1, h->colmap, h->kpad, h->n1p, scale, h->w1p,
                                           fw_kc(h->ku));
    int s = post_launch();
    if (s) return s;
  }
  if (h->n_hidden == 2) {
    const long long total = (long long)h->n2p * h->n1p;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 2048u) blocks = 2048u;
    fw_pack_kernel<<<blocks, 256, 0, st>>>(p->W[1], h->n1, h->n2, nullptr, h->n1p, h->n2p, scale, h->w2p, fw_kc(h->ku));
    int s = post_launch();
    if (s) return s;
  }
  for (int k = 0; k < p->n_layers; ++k) {
    if (h->gt_fwd[k] == nullptr) continue;
    const int K = p->dims[k], N = p->dims[k + 1];
    int s = launch_gemm_pack(p->W[k], (long long)K, 1, N, K, h->gt_fwd[k], st);          // B[n][k] = W[n, k]
    if (s) return s;
    s = launch_gemm_pack(p->W[k], 1, (long long)K, K, N, h->gt_bwd[k], st);              // B[j][c] = W[c, j]
    if (s) return s;
  }
  const int last = p->n_layers - 1;
  fw_pack_small_kernel<<<8, 256, 0, st>>>(p->b[0], h->n1, h->n1p, scale, h->b1s,
                                         h->n_hidden == 2 ? p->b[1] : nullptr, h->n2, h->n2p, scale, h->b2s,
                                         p->W[last], h->kout, h->nlast, h->nlastp, h->w3, p->b[last], h->b3);
  return post_launch();
}

bool prepared_matches(const MolannPrepared* h, const MolannPlan* p) {
  if (!h || h->magic != kPreparedMagic || !p) return false;
  if (h->n_inp != p->n_inp || h->n_align != p->n_align || h->n_entries != p->n_entries || h->d_feat != p->d_feat ||
      h->n_layers != p->n_layers || h->act != p->act_id)
    return false;
  for (int k = 0; k <= p->n_layers; ++k)
    if (h->dims[k] != p->dims[k]) return false;
  return true;
}

struct WideChoice {
  bool ok = false;
  FwParams P;
  long long grid = 0;
};

// shared-memory layout and scratch geometry for L frames.  MOLANN_B200_WIDE_STAGES / _RING / _SLOTS override the
// operand-ring depth, the frame-ring depth and the scratch sub-tiles per CTA.
WideChoice choose_wide(const MolannPrepared* h, const MolannPlan* p, long long L, const DeviceInfo& dev) {
  WideChoice ch;
  FwParams& P = ch.P;
  std::memset(&P, 0, sizeof(P));
  P.pos_atom = h->pos_atom; P.align_idx = p->align_idx; P.ref_x = p->ref_x; P.inv_ent = h->inv_ent;
  P.n_inp = p->n_inp; P.n_align = p->n_align; P.n_pos = h->n_pos; P.n_inv_ent = h->n_inv_ent; P.n_inv = h->n_inv;
  P.n_units = h->n_units; P.use_angle = p->use_angle_value; P.pos_is_align = h->pos_is_align;
  P.n_hidden = h->n_hidden; P.nkc1 = h->nkc1; P.n1p = h->n1p; P.nkc2 = h->nkc2; P.n2p = h->n2p;
  P.nlastp = h->nlastp; P.kout = h->kout;
  P.n1 = h->n1; P.n2 = h->n2;
  P.w1p = h->w1p; P.w2p = h->w2p; P.b1s = h->b1s; P.b2s = h->b2s; P.w3 = h->w3; P.b3 = h->b3;
  P.row_floats = fw_row_floats(h->nkc1, h->ku);
  P.slot_floats = FW_SUB * P.row_floats;
  // six sub-tiles per CTA keep the scratch inside L2 for C3-class rows (3.4 KB: 96 MB for the chip); rows that cannot stay
  // resident anyway (C5: 8.2 KB) take all eight so the geometry role runs further ahead of the converter (48 -> 51 M frames/s)
  const long long six_slots_bytes = 6LL * P.slot_floats * 4 * dev.sm_count;
  int slots = env_int("MOLANN_B200_WIDE_SLOTS", six_slots_bytes > (110LL << 20) ? 8 : 6);
  if (slots < 5) slots = 5;                       // a tile (4 sub-tiles) + at least one the geometry can run ahead in
  if (slots > FW_MAX_SLOTS) slots = FW_MAX_SLOTS;
  P.n_slots = slots;
  P.cta_floats = (long long)slots * P.slot_floats;
  P.ring_slot_bytes = round_up(12 * p->n_inp + 32, 128);
  Carver c;
  c.take((int)sizeof(FwBars), 16);
  P.off_b1 = c.take(h->n1p * 4, 16);
  P.off_b2 = c.take((h->n_hidden == 2 ? h->n2p : 1) * 4, 16);
  P.off_w3 = c.take(h->kout * h->nlastp * 4, 16);
  P.off_ypart = c.take(4 * FW_M * h->kout * 4, 16);
  // MOLANN_B200_WIDE_CDEPTH: raw K-chunks the converter keeps in flight
  int depth = env_int("MOLANN_B200_WIDE_CDEPTH", 2);
  if (depth < 2) depth = 2;
  if (depth > FW_MAX_CDEPTH) depth = FW_MAX_CDEPTH;
  P.conv_depth = depth;
  P.off_cstage = c.take(depth * fw_conv_chunk(h->ku), 128);
  const int fixed = c.off;
  // operand stages and the frame ring share what is left; plan tables move in when there is room
  int stages = env_int("MOLANN_B200_WIDE_STAGES", 2);
  if (stages < 2) stages = 2;
  if (stages > FW_MAX_STAGES) stages = FW_MAX_STAGES;
  int ring = env_int("MOLANN_B200_WIDE_RING", 4);
  if (ring > FW_MAX_RING) ring = FW_MAX_RING;
  const int budget = dev.max_smem_optin;
  const int FW_STAGE_BYTES = fw_stage_bytes(h->ku);
  auto need = [&](int st, int rg) { return round_up(fixed, 1024) + st * FW_STAGE_BYTES + rg * P.ring_slot_bytes; };
  while (ring > 2 && need(stages, ring) > budget) --ring;
  while (stages > 2 && need(stages, ring) > budget) --stages;
  if (ring < 2 || need(stages, ring) > budget) return ch;

  P.n_stages = stages;
  P.n_ring = ring;
  P.off_stage = c.take(stages * FW_STAGE_BYTES, 1024);
  P.off_ring = c.take(ring * P.ring_slot_bytes, 128);
  P.off_pos = P.off_aidx = P.off_ref = P.off_ent = -1;
  const bool stage_tables = env_int("MOLANN_B200_WIDE_TABLES", 1) != 0;     // 0: tests force the global-memory tables
  if (stage_tables && c.off + h->n_inv_ent * ENTRY_INTS * 4 + 16 <= budget && h->n_inv_ent > 0)
    P.off_ent = c.take(h->n_inv_ent * ENTRY_INTS * 4, 16);
  if (stage_tables && c.off + p->n_align * 16 + 32 <= budget && p->n_align > 0) {
    P.off_aidx = c.take(p->n_align * 4, 16);
    P.off_ref = c.take(p->n_align * 12, 16);
  }
  if (stage_tables && c.off + h->n_pos * 4 + 16 <= budget && h->n_pos > 0) P.off_pos = c.take(h->n_pos * 4, 16);
  P.total_smem = round_up(c.off, 128);
  if (P.total_smem > budget) return ch;
  const long long ntiles = (L + FW_M - 1) / FW_M;
  ch.grid = dev.sm_count < ntiles ? dev.sm_count : ntiles;
  if (ch.grid < 1) ch.grid = 1;
  ch.ok = true;
  return ch;
}

template <int ACT, int KU, bool STORE_H>
int launch_wide_act(const WideChoice& ch, const float* x, float* y, long long L, cudaStream_t st, float* h1, float* h2) {
  auto kern = fused_wide_forward_kernel<ACT, KU, STORE_H>;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ch.P.total_smem));
  if (s) return s;
  FwParams P = ch.P;
  P.x_base = x;
  P.h1_out = h1;
  P.h2_out = h2;
  kern<<<(unsigned)ch.grid, FW_THREADS, P.total_smem, st>>>(P, x, y, L);
  return post_launch();
}

// The activation-storing variant exists for tanh only (the reference's default and every BASELINE config); the
// value-and-gradient of other activations takes the layered route.
bool wide_can_store_h(int act_id) { return act_id == MOLANN_ACT_TANH; }

int launch_wide(const WideChoice& ch, int ku, int act_id, const float* x, float* y, long long L, cudaStream_t st,
                float* h1, float* h2) {
  if (h1 != nullptr) {
    if (!wide_can_store_h(act_id)) return MOLANN_ERR_PLAN;
    return ku == 4 ? launch_wide_act<ACT_TANH, 4, true>(ch, x, y, L, st, h1, h2)
                   : launch_wide_act<ACT_TANH, 2, true>(ch, x, y, L, st, h1, h2);
  }
#define WIDE_LAUNCH(A) \
  (ku == 4 ? launch_wide_act<A, 4, false>(ch, x, y, L, st, nullptr, nullptr) \
           : launch_wide_act<A, 2, false>(ch, x, y, L, st, nullptr, nullptr))
  switch (act_id) {
    case MOLANN_ACT_TANH: return WIDE_LAUNCH(ACT_TANH);
    case MOLANN_ACT_RELU: return WIDE_LAUNCH(ACT_RELU);
    case MOLANN_ACT_SIGMOID: return WIDE_LAUNCH(ACT_SIGMOID);
    default: return WIDE_LAUNCH(ACT_IDENTITY);
  }
#undef WIDE_LAUNCH
}

// Jacobian mode of the same kernel (one tile per CTA: five 64-column TMEM blocks per tile)
template <int ACT>
int launch_tc_jac(const DevPlan& dp, const TcVgLayout& vl, const float* x, float* y, float* jac, long long L,
                  const DeviceInfo& dev, cudaStream_t st) {
  auto kern = fused_tc_value_grad_kernel<1, ACT, true>;
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, vl.total_bytes));
  if (s) return s;
  const long long ntiles = (L + TC_F - 1) / TC_F;
  long long grid = dev.sm_count;                        // 512 TMEM columns per CTA: one CTA per SM
  if (grid > ntiles) grid = ntiles;
  const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(jac) & 15u) == 0) &&
                      ((TC_F * 3 * dp.n_inp * 4) % 16 == 0) && ((L * 3 * dp.n_inp * 4) % 16 == 0);
  kern<<<(unsigned)grid, TC_F, vl.total_bytes, st>>>(dp, vl, x, nullptr, y, jac, L, use_tma);
  return post_launch();
}

// Wire format of the ingestion path: coordinates as int16 steps of `res` around a batch origin,
//   x[f, a, c] = origin[c] + q[f, a, c] * res        (one FMA in fp32 -- the host decoder does the same, bit for bit)
// Eight values per thread: one 16-byte load, two 16-byte stores.
__global__ void __launch_bounds__(256)
decode_frames_i16_kernel(const int16_t* __restrict__ q, long long n, float o0, float o1, float o2, float res,
                         float* __restrict__ x, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n8 = n >> 3;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += stride) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(q) + t);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      int c = (int)((t * 8) % 3);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int16_t s16 = (int16_t)((w[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
        const float o = c == 0 ? o0 : (c == 1 ? o1 : o2);
        v[j] = fmaf((float)s16, res, o);
        c = c == 2 ? 0 : c + 1;
      }
      float4* dst = reinterpret_cast<float4*>(x) + 2 * t;
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int c = (int)(i % 3);
      x[i] = fmaf((float)q[i], res, c == 0 ? o0 : (c == 1 ? o1 : o2));
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int c = (int)(i % 3);
      x[i] = fmaf((float)q[i], res, c == 0 ? o0 : (c == 1 ? o1 : o2));
    }
  }
}

__global__ void onehot_rows_kernel(float* __restrict__ cot, long long L, int k, int plane) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L * k; i += (long long)gridDim.x * blockDim.x)
    cot[i] = (int)(i % k) == plane ? 1.f : 0.f;
}

#define SMALL_DISPATCH(ch, FN, ...)                                   \
  ((ch).F == 128 ? ((ch).NT == 256 ? FN<128, 256>(__VA_ARGS__) : FN<128, 128>(__VA_ARGS__)) \
                 : ((ch).NT == 256 ? FN<64, 256>(__VA_ARGS__) : FN<64, 128>(__VA_ARGS__)))

// ---------------------------------------------------------------------------------------------
// general path
// ---------------------------------------------------------------------------------------------
constexpr long long kChunkFrames = 32768;

// Frames per pass of the layered path: scratch stays O(chunk).  32768 frames for wide systems (C3: 11 KB of
// activations per frame); narrow ones (C2 / C4: 1.2 KB) take up to 8x that within the same 256 MB, which cuts the
// launches of a 2^20-frame training step from 482 to ~90.  MOLANN_B200_CHUNK overrides.
long long chunk_frames(const MolannPlan* p, long long L) {
  long long per_frame = p->d_feat;
  int widest = p->d_feat;
  for (int k = 1; k <= p->n_layers; ++k) {
    per_frame += p->dims[k];
    widest = p->dims[k] > widest ? p->dims[k] : widest;
  }
  per_frame = 4 * (per_frame + 2LL * widest);
  long long ch = (256LL << 20) / (per_frame > 0 ? per_frame : 1) / kChunkFrames * kChunkFrames;
  if (ch < kChunkFrames) ch = kChunkFrames;
  if (ch > 8 * kChunkFrames) ch = 8 * kChunkFrames;
  ch = env_int("MOLANN_B200_CHUNK", (int)ch);
  if (ch < 256) ch = 256;
  return L < ch ? L : ch;
}

int max_dim(const MolannPlan* p) {
  int m = p->d_feat;
  for (int k = 0; k <= p->n_layers; ++k) m = p->dims[k] > m ? p->dims[k] : m;
  return m;
}


// scratch for the packed (TF32 hi/lo, chunk-major) weights of ONE tensor-core GEMM; packed right before each use
size_t gemm_pack_bytes(const MolannPlan* p) {
  long long m = 0;
  for (int k = 0; k < p->n_layers; ++k) {
    const long long f = gt_pack_floats(p->dims[k + 1], p->dims[k]), b = gt_pack_floats(p->dims[k], p->dims[k + 1]);
    m = f > m ? f : m;
    m = b > m ? b : m;
  }
  return align256((size_t)m * 4);
}

size_t general_ws_bytes(const MolannPlan* p, long long L, bool backward) {
  const long long ch = chunk_frames(p, L);
  size_t total = align256((size_t)ch * p->d_feat * 4) + gemm_pack_bytes(p);
  if (!backward) {
    total += 2 * align256((size_t)ch * max_dim(p) * 4);
  } else {
    for (int k = 1; k < p->n_layers; ++k) total += align256((size_t)ch * p->dims[k] * 4);
    total += 2 * align256((size_t)ch * max_dim(p) * 4);
  }
  return total;
}

// Staged preprocess kernels (general.cuh): every warp owns a frame buffer (two for the backward) in shared memory.
// MOLANN_B200_STAGED = 0 keeps the gather kernels.
struct StagedChoice {
  bool ok = false;
  int warps = 0, buf_bytes = 0, fbuf_bytes = 0, smem = 0;
};
StagedChoice choose_staged(const MolannPlan* p, bool backward, const DeviceInfo& dev) {
  StagedChoice ch;
  if (env_int("MOLANN_B200_STAGED", 1) == 0 || p->n_inp < 256) return ch;     // small frames: gathers are fine
  ch.buf_bytes = round_up(3 * p->n_inp * 4 + 32, 128);
  ch.fbuf_bytes = round_up(p->d_feat * 4 + 32, 128);
  const int per_warp = backward ? 2 * ch.buf_bytes + ch.fbuf_bytes : ch.buf_bytes;
  int w = (dev.max_smem_optin - 128) / per_warp;
  if (w > WARPS_PER_CTA) w = WARPS_PER_CTA;
  if (w < 2) return ch;
  ch.warps = w;
  ch.smem = 128 + w * per_warp;
  ch.ok = true;
  return ch;
}
// Block-per-frame pipelined kernels (staged_block.cuh): MOLANN_B200_STAGED = 2 (default) for big frames; = 1 keeps
// the warp-per-frame staged kernels, = 0 the gather kernels.  Ring depth: MOLANN_B200_SB_STAGES (default 2, the
// minimum), reduced towards 2 while it costs the second CTA per SM.
struct BlockChoice {
  bool ok = false;
  SbLayout lay;
};
BlockChoice choose_block(const MolannPlan* p, bool backward, const DeviceInfo& dev) {
  BlockChoice ch;
  const int mode = env_int("MOLANN_B200_STAGED", 2);
  if (mode < 2 || p->n_inp < 256) return ch;
  // forward: the warp-per-frame staged kernel still wins while 5+ frames fit on an SM (C3: 0.20 vs 0.23 ms per 32768
  // frames); the role pipeline takes over for frames beyond 40 KB (C5: 0.36 vs 0.38 ms).  MOLANN_B200_STAGED = 3
  // forces the block kernel in both directions.
  if (!backward && mode < 3 && 3 * p->n_inp * 4 <= 40 * 1024) return ch;
  SbLayout& lay = ch.lay;
  lay.buf_bytes = round_up(3 * p->n_inp * 4 + 32, 128);
  lay.fbuf_bytes = backward ? round_up(p->d_feat * 4 + 32, 128) : 0;
  int off = SB_HEAD;
  lay.aidx_off = off; off += round_up(4 * p->n_align, 16);
  lay.ref_off = off; off += round_up(12 * p->n_align, 16);
  lay.ent_off = off; off += round_up(4 * MOLANN_ENTRY_INTS * p->n_entries, 128);
  lay.zero_off = -1;
  if (backward && env_int("MOLANN_B200_SB_BULKZERO", 1) != 0) {
    lay.zero_off = off;
    off += SB_ZERO_BYTES;
  }
  lay.ring_off = off;
  const int slot = lay.buf_bytes + lay.fbuf_bytes;
  int stages = env_int("MOLANN_B200_SB_STAGES", 2);
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;                                 // the roles run one frame apart: two slots at least
  const int two_ctas = (dev.smem_per_sm / 2) - 1024;          // budget per CTA that keeps two resident
  while (stages > 2 && off + stages * slot > two_ctas) --stages;
  if (off + stages * slot > dev.max_smem_optin) return ch;
  lay.stages = stages;
  lay.total = off + stages * slot;
  ch.ok = true;
  return ch;
}
template <class Kern>
int block_grid(Kern kern, int smem, long long L, const DeviceInfo& dev, long long* grid) {
  int s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (s) return s;
  int per_sm = 0;
  s = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SB_THREADS, (size_t)smem));
  if (s) return s;
  const int want = env_int("MOLANN_B200_SB_CTAS", 0);
  if (want > 0 && want < per_sm) per_sm = want;
  *grid = (long long)dev.sm_count * (per_sm < 1 ? 1 : per_sm);
  if (*grid > L) *grid = L;
  return MOLANN_OK;
}
// Thread-per-frame tile kernels (small_tile.cuh) for frames small enough that 128 of them (+ the result tile) fit in
// shared memory twice over (two CTAs per SM).  MOLANN_B200_TILE = 0 keeps the warp-per-frame kernels.
struct TileChoice {
  bool ok = false;
  StLayout lay;
};
TileChoice choose_tile(const MolannPlan* p, bool backward, const DeviceInfo& dev) {
  TileChoice ch;
  if (env_int("MOLANN_B200_TILE", 1) == 0 || p->n_entries < 1) return ch;
  StLayout& lay = ch.lay;
  int off = 0;
  lay.aidx_off = off; off += round_up(4 * p->n_align, 16);
  lay.ref_off = off; off += round_up(12 * p->n_align, 16);
  lay.ent_off = off; off += round_up(4 * MOLANN_ENTRY_INTS * p->n_entries, 128);
  lay.xs_off = off; off += round_up(ST_F * 3 * p->n_inp * 4, 128);
  lay.out_off = off; off += round_up(ST_F * (backward ? 3 * p->n_inp : p->d_feat) * 4, 128);
  lay.gf_off = off;
  if (backward) off += round_up(ST_F * p->d_feat * 4, 128);
  lay.total = off;
  if (lay.total > (dev.smem_per_sm / 2) - 1024) return ch;
  ch.ok = true;
  return ch;
}
TileChoice choose_align_tile(const MolannPlan* p, bool backward, const DeviceInfo& dev) {
  TileChoice ch;
  if (env_int("MOLANN_B200_TILE", 1) == 0) return ch;
  StLayout& lay = ch.lay;
  int off = 0;
  lay.aidx_off = off; off += round_up(4 * p->n_align, 16);
  lay.ref_off = off; off += round_up(12 * p->n_align, 16);
  lay.ent_off = off;
  lay.xs_off = off; off += round_up(ST_F * 3 * p->n_inp * 4, 128);
  lay.out_off = off;
  if (backward) off += round_up(ST_F * 3 * p->n_inp * 4, 128);
  lay.gf_off = off;
  lay.total = off;
  if (lay.total > (dev.smem_per_sm / 2) - 1024) return ch;
  ch.ok = true;
  return ch;
}
unsigned tile_grid(long long L, int smem, const DeviceInfo& dev) {
  long long grid = (L + ST_F - 1) / ST_F;
  const long long cap = (long long)dev.sm_count * (dev.smem_per_sm / (smem + 1024));
  if (grid > cap) grid = cap;
  return (unsigned)(grid < 1 ? 1 : grid);
}
int launch_preprocess_forward(const MolannPlan* p, const DevPlan& dp, const float* x, float* feat, long long L,
                              const DeviceInfo& dev, cudaStream_t st) {
  const TileChoice tc = choose_tile(p, false, dev);
  if (tc.ok) {
    int s = check_cuda(cudaFuncSetAttribute(preprocess_forward_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            tc.lay.total));
    if (s) return s;
    long long grid = (L + ST_F - 1) / ST_F;
    const long long cap = (long long)dev.sm_count * (dev.smem_per_sm / (tc.lay.total + 1024));
    if (grid > cap) grid = cap;
    preprocess_forward_tile_kernel<<<(unsigned)grid, ST_F, tc.lay.total, st>>>(dp, tc.lay, x, feat, L);
    return post_launch();
  }
  const BlockChoice bc = choose_block(p, false, dev);
  if (bc.ok) {
    long long grid = 1;
    int s = block_grid(preprocess_forward_block_kernel, bc.lay.total, L, dev, &grid);
    if (s) return s;
    preprocess_forward_block_kernel<<<(unsigned)grid, SB_THREADS, bc.lay.total, st>>>(dp, bc.lay, x, feat, L);
    return post_launch();
  }
  const StagedChoice sc = choose_staged(p, false, dev);
  if (sc.ok) {
    int s = check_cuda(cudaFuncSetAttribute(preprocess_forward_staged_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, sc.smem));
    if (s) return s;
    long long grid = (L + sc.warps - 1) / sc.warps;
    if (grid > dev.sm_count) grid = dev.sm_count;
    preprocess_forward_staged_kernel<<<(unsigned)grid, sc.warps * 32, sc.smem, st>>>(dp, x, feat, L, sc.buf_bytes);
    return post_launch();
  }
  long long blocks = (L + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const long long cap = (long long)dev.sm_count * 8;
  if (blocks > cap) blocks = cap;
  preprocess_forward_warp_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), WARPS_PER_CTA * 32, 0, st>>>(dp, x, feat, L);
  return post_launch();
}
int launch_preprocess_backward(const MolannPlan* p, const DevPlan& dp, const float* x, const float* gfeat, float* gx,
                               long long L, const DeviceInfo& dev, cudaStream_t st) {
  const TileChoice tc = choose_tile(p, true, dev);
  if (tc.ok) {
    int s = check_cuda(cudaFuncSetAttribute(preprocess_backward_tile_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, tc.lay.total));
    if (s) return s;
    long long grid = (L + ST_F - 1) / ST_F;
    const long long cap = (long long)dev.sm_count * (dev.smem_per_sm / (tc.lay.total + 1024));
    if (grid > cap) grid = cap;
    preprocess_backward_tile_kernel<<<(unsigned)grid, ST_F, tc.lay.total, st>>>(dp, tc.lay, x, gfeat, gx, L);
    return post_launch();
  }
  const BlockChoice bc = choose_block(p, true, dev);
  if (bc.ok) {
    long long grid = 1;
    int s = block_grid(preprocess_backward_block_kernel, bc.lay.total, L, dev, &grid);
    if (s) return s;
    preprocess_backward_block_kernel<<<(unsigned)grid, SB_THREADS, bc.lay.total, st>>>(dp, bc.lay, x, gfeat, gx, L);
    return post_launch();
  }
  const StagedChoice sc = choose_staged(p, true, dev);
  if (sc.ok) {
    int s = check_cuda(cudaFuncSetAttribute(preprocess_backward_staged_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, sc.smem));
    if (s) return s;
    long long grid = (L + sc.warps - 1) / sc.warps;
    if (grid > dev.sm_count) grid = dev.sm_count;
    preprocess_backward_staged_kernel<<<(unsigned)grid, sc.warps * 32, sc.smem, st>>>(dp, x, gfeat, gx, L,
                                                                                      sc.buf_bytes, sc.fbuf_bytes);
    return post_launch();
  }
  long long blocks = (L + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const long long cap = (long long)dev.sm_count * 8;
  if (blocks > cap) blocks = cap;
  preprocess_backward_warp_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), WARPS_PER_CTA * 32, 0, st>>>(dp, x, gfeat, gx,
                                                                                                   L);
  return post_launch();
}

unsigned warp_grid(long long L, const DeviceInfo& dev) {
  long long blocks = (L + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const long long cap = (long long)dev.sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// Tensor-core GEMM (gemm_tc.cuh) for layers wide enough to fill its 128 x N x 32 tiles; MOLANN_B200_GEMM_TC = 0
// keeps the FFMA kernel (A/B testing).
bool use_gemm_tc(long long M, int K, int N, const void* pack) {
  if (pack == nullptr || env_int("MOLANN_B200_GEMM_TC", 1) == 0 || env_int("MOLANN_B200_TC", 1) == 0) return false;
  return M >= 64 && K >= 32 && N >= 32;
}

template <int EPI>
int launch_gemm_tc(const float* A, long long M, int K, const float* W, long long rs, long long cs, int N, float* C,
                   const float* bias, const float* H, int act, int apply_act, float* pack, const DeviceInfo& dev,
                   cudaStream_t st, const float* prepacked = nullptr) {
  int s = MOLANN_OK;
  if (prepacked == nullptr) {                  // no prepared plan: pack into the workspace right before the use
    s = launch_gemm_pack(W, rs, cs, N, K, pack, st);
    if (s) return s;
  } else {
    pack = const_cast<float*>(prepacked);
  }
  auto kern = gemm_tc_kernel<EPI>;
  s = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES));
  if (s) return s;
  const long long items = ((M + GT_M - 1) / GT_M) * ((N + GT_NMAX - 1) / GT_NMAX);
  long long grid = dev.sm_count;
  if (grid > items) grid = items;
  int seg = env_int("MOLANN_B200_GEMM_SEG", GT_SEG);
  if (seg < 1) seg = 1;
  kern<<<(unsigned)grid, GT_THREADS, GT_SMEM_BYTES, st>>>(A, (long long)K, M, K, pack, N, C, (long long)N, bias, H, act,
                                                          apply_act, seg);
  return post_launch();
}

int launch_linear_forward(const float* in, const float* W, const float* b, float* out, long long M, int K, int N,
                          int act, int apply_act, cudaStream_t st, float* pack = nullptr,
                          const DeviceInfo* dev = nullptr, const float* prepacked = nullptr) {
  if (dev != nullptr && use_gemm_tc(M, K, N, pack))
    return launch_gemm_tc<GT_EPI_BIAS_ACT>(in, M, K, W, (long long)K, 1, N, out, b, nullptr, act, apply_act, pack, *dev,
                                           st, prepacked);
  if (N <= NARROW_MAX && env_int("MOLANN_B200_NARROW", 1) != 0) {
    long long blocks = (M + 7) / 8;
    const long long cap = (long long)(dev ? dev->sm_count : 148) * 8;
    if (blocks > cap) blocks = cap;
    narrow_forward_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(in, W, b, out, M, K, N, act, apply_act);
    return post_launch();
  }
  dim3 grid((N + 63) / 64, (unsigned)((M + 63) / 64), 1);
  gemm_kernel<EPI_BIAS_ACT><<<grid, 256, 0, st>>>(in, K, 1, W, 1, K, out, N, (int)M, N, K, K, b, nullptr, act,
                                                  apply_act);
  return post_launch();
}

// gprev[M, K] = (gz[M, N] W[N, K]) * act'(hprev[M, K])
int launch_linear_backward_input(const float* gz, const float* W, const float* hprev, float* gprev, long long M,
                                 int K, int N, int act, cudaStream_t st, float* pack = nullptr,
                                 const DeviceInfo* dev = nullptr, const float* prepacked = nullptr) {
  // gprev[M x K] = gz[M x N] * W[N x K]: as C = A B^T with B[j][c] = W[c * K + j] (rows = inputs, contraction = outputs)
  if (dev != nullptr && use_gemm_tc(M, N, K, pack))
    return launch_gemm_tc<GT_EPI_DACT>(gz, M, N, W, 1, (long long)K, K, gprev, nullptr, hprev, act, hprev != nullptr, pack,
                                       *dev, st, prepacked);
  if (N <= NARROW_MAX && env_int("MOLANN_B200_NARROW", 1) != 0) {
    long long blocks = (M * ((K + 3) / 4) + 255) / 256;
    const long long cap = (long long)(dev ? dev->sm_count : 148) * 16;
    if (blocks > cap) blocks = cap;
    narrow_backward_input_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(gz, W, hprev, gprev, M, K, N,
                                                                                     act);
    return post_launch();
  }
  dim3 grid((K + 63) / 64, (unsigned)((M + 63) / 64), 1);
  gemm_kernel<EPI_DACT><<<grid, 256, 0, st>>>(gz, N, 1, W, K, 1, gprev, K, (int)M, K, N, N, nullptr, hprev, act,
                                              hprev != nullptr);
  return post_launch();
}

// gW[N, K] += gz[M, N]^T hin[M, K];  gb[N] += colsum(gz)
// The contraction runs over the FRAMES, the output is tiny (64 x 64 for C2/C4), so the frame axis is what fills the
// GPU: it is cut so that about two CTAs per SM exist (a fixed 2048-row cut left 16 CTAs for a 32768-frame chunk and
// made this kernel 60 % of a C4 training step); partial results meet in L2 through fp32 atomics.
int launch_linear_backward_params(const float* gz, const float* hin, float* gW, float* gb, long long M, int K, int N,
                                  cudaStream_t st, int sm_count = 148) {
  const long long tiles = (long long)((K + 63) / 64) * ((N + 63) / 64);
  long long slabs = (2LL * sm_count + tiles - 1) / tiles;
  if (slabs < 1) slabs = 1;
  long long kchunk = (M + slabs - 1) / slabs;
  kchunk = (kchunk + 15) / 16 * 16;
  if (kchunk < 64) kchunk = 64;
  if (kchunk > 2048) kchunk = 2048;
  dim3 grid((K + 63) / 64, (N + 63) / 64, (unsigned)((M + kchunk - 1) / kchunk));
  gemm_kernel<EPI_ATOMIC><<<grid, 256, 0, st>>>(gz, 1, N, hin, K, 1, gW, K, N, K, M, kchunk, nullptr, nullptr, 0, 0);
  int s = post_launch();
  if (s) return s;
  if (gb) {
    const long long cols = (N + 31) / 32;
    long long rslabs = (2LL * sm_count + cols - 1) / cols;
    long long rows = (M + rslabs - 1) / rslabs;
    rows = (rows + 7) / 8 * 8;
    if (rows < 64) rows = 64;
    if (rows > 4096) rows = 4096;
    dim3 g2((unsigned)cols, (unsigned)((M + rows - 1) / rows), 1);
    colsum_atomic_kernel<<<g2, 256, 0, st>>>(gz, (int)M, N, gb, (int)rows);
    s = post_launch();
  }
  return s;
}

int general_forward(const MolannPlan* p, const float* x, long long L, float* y, void* ws, size_t ws_bytes,
                    const DeviceInfo& dev, cudaStream_t st, const MolannPrepared* prep = nullptr) {
  if (!ws || ws_bytes < general_ws_bytes(p, L, false)) return MOLANN_ERR_WORKSPACE;
  const DevPlan dp = to_dev(p);
  const long long ch = chunk_frames(p, L);
  char* base = static_cast<char*>(ws);
  float* feat = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * p->d_feat * 4);
  float* pack = reinterpret_cast<float*>(base);
  base += gemm_pack_bytes(p);
  float* pp[2];
  pp[0] = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * max_dim(p) * 4);
  pp[1] = reinterpret_cast<float*>(base);
  const int kout = p->dims[p->n_layers];
  for (long long c0 = 0; c0 < L; c0 += ch) {
    const long long Lc = (L - c0 < ch) ? (L - c0) : ch;
    int s = launch_preprocess_forward(p, dp, x + c0 * 3 * p->n_inp, feat, Lc, dev, st);
    if (s) return s;
    const float* in = feat;
    for (int k = 0; k < p->n_layers; ++k) {
      const bool last = (k == p->n_layers - 1);
      float* out = last ? (y + c0 * kout) : pp[k & 1];
      s = launch_linear_forward(in, p->W[k], p->b[k], out, Lc, p->dims[k], p->dims[k + 1], p->act_id, !last, st, pack,
                                &dev, prep ? prep->gt_fwd[k] : nullptr);
      if (s) return s;
      in = out;
    }
  }
  return MOLANN_OK;
}

// y_out != nullptr: also write the model outputs (value_and_grad: the forward the backward recomputes anyway is the
// forward the caller wanted -- running molann_b200_forward first cost a second pass over x and both wide GEMMs).
int general_backward(const MolannPlan* p, const float* x, const float* gy, long long L, float* gx, float* const* gW,
                     float* const* gb, void* ws, size_t ws_bytes, const DeviceInfo& dev, cudaStream_t st,
                     float* y_out = nullptr, const MolannPrepared* prep = nullptr) {
  if (!ws || ws_bytes < general_ws_bytes(p, L, true)) return MOLANN_ERR_WORKSPACE;
  const DevPlan dp = to_dev(p);
  const long long ch = chunk_frames(p, L);
  const int nl = p->n_layers;
  char* base = static_cast<char*>(ws);
  float* h[MOLANN_MAX_LAYERS];       // h[0] = features, h[k] = activation after layer k
  h[0] = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * p->d_feat * 4);
  float* pack = reinterpret_cast<float*>(base);
  base += gemm_pack_bytes(p);
  for (int k = 1; k < nl; ++k) {
    h[k] = reinterpret_cast<float*>(base);
    base += align256((size_t)ch * p->dims[k] * 4);
  }
  float* pp[2];
  pp[0] = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * max_dim(p) * 4);
  pp[1] = reinterpret_cast<float*>(base);
  const int kout = p->dims[nl];
  for (long long c0 = 0; c0 < L; c0 += ch) {
    const long long Lc = (L - c0 < ch) ? (L - c0) : ch;
    const float* xc = x + c0 * 3 * p->n_inp;
    int s = launch_preprocess_forward(p, dp, xc, h[0], Lc, dev, st);
    if (s) return s;
    for (int k = 0; k < nl - 1; ++k) {
      s = launch_linear_forward(h[k], p->W[k], p->b[k], h[k + 1], Lc, p->dims[k], p->dims[k + 1], p->act_id, 1, st,
                                pack, &dev, prep ? prep->gt_fwd[k] : nullptr);
      if (s) return s;
    }
    if (y_out != nullptr) {
      s = launch_linear_forward(h[nl - 1], p->W[nl - 1], p->b[nl - 1], y_out + c0 * kout, Lc, p->dims[nl - 1], kout,
                                p->act_id, 0, st, pack, &dev, prep ? prep->gt_fwd[nl - 1] : nullptr);
      if (s) return s;
    }
    const float* gz = gy + c0 * kout;
    for (int k = nl - 1; k >= 0; --k) {
      if (gW && gW[k]) {
        s = launch_linear_backward_params(gz, h[k], gW[k], gb ? gb[k] : nullptr, Lc, p->dims[k], p->dims[k + 1], st,
                                          dev.sm_count);
        if (s) return s;
      }
      if (k == 0 && gx == nullptr) break;       // training: nobody asked for the coordinate gradient
      float* gprev = pp[k & 1];
      s = launch_linear_backward_input(gz, p->W[k], k > 0 ? h[k] : nullptr, gprev, Lc, p->dims[k], p->dims[k + 1],
                                       p->act_id, st, pack, &dev, prep ? prep->gt_bwd[k] : nullptr);
      if (s) return s;
      gz = gprev;
    }
    if (gx == nullptr) continue;
    s = launch_preprocess_backward(p, dp, xc, gz, gx + c0 * 3 * p->n_inp, Lc, dev, st);
    if (s) return s;
  }
  return MOLANN_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int molann_b200_version(void) { return MOLANN_B200_VERSION; }

const char* molann_b200_strerror(int status) {
  switch (status) {
    case MOLANN_OK: return "ok";
    case MOLANN_ERR_NULL: return "required pointer is NULL";
    case MOLANN_ERR_PLAN: return "inconsistent plan";
    case MOLANN_ERR_WORKSPACE: return "workspace missing or too small";
    case MOLANN_ERR_ALIGNMENT: return "pointer is not 4-byte aligned";
    case MOLANN_ERR_CUDA: return "CUDA runtime error (see molann_b200_last_cuda_error)";
    case MOLANN_ERR_UNSUPPORTED: return "request not supported for this plan";
    default: return "unknown status";
  }
}

int molann_b200_last_cuda_error(void) { return t_last_cuda_error; }

const char* molann_b200_cuda_error_string(int cuda_error) {
  return cudaGetErrorString(static_cast<cudaError_t>(cuda_error));
}

int64_t molann_b200_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int molann_b200_plan_validate(const MolannPlan* plan) { return validate_full(plan); }

int molann_b200_path_for(const MolannPlan* plan, int want_backward) {
  const int fam = molann_b200_kernel_family(plan, want_backward);
  return fam < 0 ? fam : (fam > 0 ? 1 : 0);
}

int molann_b200_kernel_family(const MolannPlan* plan, int want_backward) {
  if (validate_full(plan) != MOLANN_OK) return -1;
  DeviceInfo dev = device_info();
  if (!dev.ok) { dev.sm_count = 148; dev.max_smem_optin = 232448; dev.smem_per_sm = 233472; }   // B200 figures
  if (want_backward ? choose_tc_vg(plan, dev).ok : choose_tc(plan, false, dev).ok) return 2;
  return choose_small(plan, want_backward != 0, dev).ok ? 1 : 0;
}

size_t molann_b200_workspace_bytes(const MolannPlan* plan, int64_t L, int want_backward) {
  if (validate_full(plan) != MOLANN_OK || plan->n_layers < 1 || L <= 0) return 0;
  // always sized for the general path: a backward with parameter gradients uses it even when the
  // fused kernel serves the forward
  return general_ws_bytes(plan, L, want_backward != 0);
}

int molann_b200_forward(const MolannPlan* plan, const float* x, int64_t L, float* y, void* workspace,
                        size_t workspace_bytes, void* stream) {
  int s = validate_full(plan);
  if (s) return s;
  if (plan->n_layers < 1) return MOLANN_ERR_UNSUPPORTED;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !y) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(y)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WsChoice ws = choose_ws(plan, x, dev);
  if (ws.ok) return launch_ws_forward(to_dev(plan), ws.wl, x, y, (long long)L, dev, st);
  const TcChoice tc = choose_tc(plan, false, dev);
  if (tc.ok) return launch_tc_forward(to_dev(plan), tc.lay, x, y, (long long)L, dev, st);
  const SmallChoice ch = choose_small(plan, false, dev);
  if (ch.ok) {
    const DevPlan dp = to_dev(plan);
    return SMALL_DISPATCH(ch, launch_small_forward, dp, ch.lay, x, y, (long long)L, dev, st);
  }
  return general_forward(plan, x, L, y, workspace, workspace_bytes, dev, st);
}

int molann_b200_backward(const MolannPlan* plan, const float* x, const float* gy, int64_t L, float* gx,
                         float* const* gW, float* const* gb, void* workspace, size_t workspace_bytes,
                         void* stream) {
  int s = validate_full(plan);
  if (s) return s;
  if (plan->n_layers < 1) return MOLANN_ERR_UNSUPPORTED;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  bool want_params = false;
  if (gW)
    for (int k = 0; k < plan->n_layers; ++k) want_params = want_params || (gW[k] != nullptr);
  if (!x || !gy || (!gx && !want_params)) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(gy) || misaligned4(gx)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!want_params) {
    const TcVgChoice vg = choose_tc_vg(plan, dev);
    if (vg.ok) return run_tc_vg(vg, to_dev(plan), x, gy, nullptr, gx, (long long)L, dev, st);
    const SmallChoice ch = choose_small(plan, true, dev);
    if (ch.ok) {
      const DevPlan dp = to_dev(plan);
      return SMALL_DISPATCH(ch, launch_small_backward, dp, ch.lay, x, gy, gx, (long long)L, dev, st);
    }
  }
  return general_backward(plan, x, gy, L, gx, gW, gb, workspace, workspace_bytes, dev, st);
}

int molann_b200_value_and_grad(const MolannPlan* plan, const float* x, const float* gy, int64_t L, float* y,
                               float* gx, void* workspace, size_t workspace_bytes, void* stream) {
  int s = validate_full(plan);
  if (s) return s;
  if (plan->n_layers < 1) return MOLANN_ERR_UNSUPPORTED;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !gy || !y || !gx) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(gy) || misaligned4(y) || misaligned4(gx)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TcVgChoice vg = choose_tc_vg(plan, dev);
  if (vg.ok) return run_tc_vg(vg, to_dev(plan), x, gy, y, gx, (long long)L, dev, st);   // ONE kernel
  const SmallChoice ch = choose_small(plan, true, dev);
  if (ch.ok) {                                   // fused FFMA family: two launches (forward, backward)
    s = molann_b200_forward(plan, x, L, y, workspace, workspace_bytes, stream);
    if (s) return s;
    return molann_b200_backward(plan, x, gy, L, gx, nullptr, nullptr, workspace, workspace_bytes, stream);
  }
  // layered path: ONE pass -- the backward's forward recompute also writes y
  return general_backward(plan, x, gy, L, gx, nullptr, nullptr, workspace, workspace_bytes, dev, st, y);
}

// ---- prepared plans (fused wide kernel) ----
int molann_b200_wide_eligible(const MolannPlan* plan) {
  if (!wide_shape_ok(plan)) return 0;
  const int mode = env_int("MOLANN_B200_WIDE", -1);
  if (mode == 0 || env_int("MOLANN_B200_TC", 1) == 0) return 0;
  if (mode == 1) return 1;
  if (env_int("MOLANN_B200_PATH", -1) == 0) return 0;             // tests forcing the layered path
  return molann_b200_kernel_family(plan, 0) == 0 ? 1 : 0;
}

size_t molann_b200_prepared_bytes(const MolannPlan* plan) {
  if (!wide_shape_ok(plan)) return 0;
  // the program lives in device memory: bound its regrouped size from the plan's scalar fields
  const int n_units_max = plan->n_entries + (plan->d_feat + 3) / 4 + 1;
  return wide_prepared_bytes_for(plan, plan->n_entries, plan->n_entries, n_units_max);
}

int molann_b200_prepare(const MolannPlan* plan, void* device_buffer, size_t bytes, void* stream,
                        MolannPrepared** out) {
  if (!out) return MOLANN_ERR_NULL;
  *out = nullptr;
  int s = validate_full(plan);
  if (s) return s;
  if (!wide_shape_ok(plan)) return MOLANN_ERR_UNSUPPORTED;
  if (!device_buffer) return MOLANN_ERR_NULL;
  if (bytes < molann_b200_prepared_bytes(plan)) return MOLANN_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the feature program, regrouped on the host: position atoms first (one 16-byte unit each), invariant features after
  std::vector<int32_t> ent((size_t)plan->n_entries * ENTRY_INTS);
  s = check_cuda(cudaMemcpyAsync(ent.data(), plan->entries, ent.size() * 4, cudaMemcpyDeviceToHost, st));
  if (s) return s;
  s = check_cuda(cudaStreamSynchronize(st));
  if (s) return s;
  std::vector<int32_t> pos_atom, pos_col, inv_ent, inv_col;
  const int inv_dim_dihedral = plan->use_angle_value ? 1 : 2;
  for (int e = 0; e < plan->n_entries; ++e) {
    const int32_t* en = &ent[(size_t)e * ENTRY_INTS];
    if (en[0] == MOLANN_FEAT_POSITION) {
      pos_atom.push_back(en[1]);
      pos_col.push_back(en[5]);
    } else {
      const int dim = en[0] == MOLANN_FEAT_DIHEDRAL ? inv_dim_dihedral : 1;
      for (int q = 0; q < 5; ++q) inv_ent.push_back(en[q]);
      inv_ent.push_back((int32_t)inv_col.size());               // first invariant column of this entry
      for (int q = 0; q < dim; ++q) inv_col.push_back(en[5] + q);
    }
    for (int q = 1; q <= 4; ++q)
      if (en[q] < 0 || en[q] >= plan->n_inp) return MOLANN_ERR_PLAN;
  }
  MolannPrepared* h = new (std::nothrow) MolannPrepared;
  if (!h) return MOLANN_ERR_NULL;
  std::memset(h, 0, sizeof(*h));
  h->magic = kPreparedMagic;
  h->n_inp = plan->n_inp; h->n_align = plan->n_align; h->n_entries = plan->n_entries; h->d_feat = plan->d_feat;
  h->n_layers = plan->n_layers; h->act = plan->act_id;
  for (int k = 0; k <= plan->n_layers; ++k) h->dims[k] = plan->dims[k];
  h->n_pos = (int)pos_atom.size();
  h->n_inv_ent = (int)inv_ent.size() / ENTRY_INTS;
  h->n_inv = (int)inv_col.size();
  h->n_units = wide_units(h->n_pos, h->n_inv);
  // position atoms == alignment selection (same order)?  Then the moments pass also delivers the position units.
  h->pos_is_align = 0;
  if (plan->n_align > 0 && plan->n_align == h->n_pos) {
    std::vector<int32_t> al((size_t)plan->n_align);
    s = check_cuda(cudaMemcpyAsync(al.data(), plan->align_idx, al.size() * 4, cudaMemcpyDeviceToHost, st));
    if (!s) s = check_cuda(cudaStreamSynchronize(st));
    if (s) { delete h; return s; }
    h->pos_is_align = 1;
    for (int u = 0; u < h->n_pos; ++u)
      if (al[u] != pos_atom[u]) { h->pos_is_align = 0; break; }
  }
  h->ku = wide_ku_for(plan);
  const int FW_KU = h->ku, FW_KC = fw_kc(h->ku);
  h->nkc1 = (h->n_units + FW_KU - 1) / FW_KU;
  if (h->nkc1 < 1) h->nkc1 = 1;
  h->kpad = h->nkc1 * FW_KC;
  h->n_hidden = plan->n_layers - 1;
  h->n1 = plan->dims[1]; h->n1p = round_up(h->n1, 16);
  h->n2 = h->n_hidden == 2 ? plan->dims[2] : 0; h->n2p = h->n_hidden == 2 ? round_up(h->n2, 16) : 16;
  h->nkc2 = h->n_hidden == 2 ? h->n1p / FW_KC : 0;
  h->nlast = h->n_hidden == 2 ? h->n2 : h->n1;
  h->nlastp = h->n_hidden == 2 ? h->n2p : h->n1p;
  h->kout = plan->dims[plan->n_layers];
  // internal K index -> original feature column (-1: padding)
  std::vector<int32_t> colmap((size_t)h->kpad, -1);
  for (int u = 0; u < h->n_pos; ++u)
    for (int q = 0; q < 3; ++q) colmap[(size_t)4 * u + q] = pos_col[u] + q;
  for (int v = 0; v < h->n_inv; ++v) colmap[(size_t)(v < h->n_pos ? 4 * v + 3 : 3 * h->n_pos + v)] = inv_col[v];
  // carve the caller's buffer
  char* base = static_cast<char*>(device_buffer);
  auto take = [&](size_t nbytes) { char* r = base; base += align256(nbytes); return r; };
  h->pos_atom = reinterpret_cast<int*>(take((size_t)h->n_pos * 4));
  h->inv_ent = reinterpret_cast<int*>(take((size_t)h->n_inv_ent * ENTRY_INTS * 4));
  h->colmap = reinterpret_cast<int*>(take((size_t)h->kpad * 4));
  h->w1p = reinterpret_cast<float*>(take((size_t)h->nkc1 * 2 * FW_KC * h->n1p * 4));
  h->b1s = reinterpret_cast<float*>(take((size_t)h->n1p * 4));
  if (h->n_hidden == 2) {
    h->w2p = reinterpret_cast<float*>(take((size_t)h->nkc2 * 2 * FW_KC * h->n2p * 4));
    h->b2s = reinterpret_cast<float*>(take((size_t)h->n2p * 4));
  }
  h->w3 = reinterpret_cast<float*>(take((size_t)h->kout * h->nlastp * 4));
  h->b3 = reinterpret_cast<float*>(take((size_t)h->kout * 4));
  for (int k = 0; k < plan->n_layers; ++k) {
    if (plan->dims[k] >= 32 && plan->dims[k + 1] >= 32) {
      h->gt_fwd[k] = reinterpret_cast<float*>(take((size_t)gt_pack_floats(plan->dims[k + 1], plan->dims[k]) * 4));
      h->gt_bwd[k] = reinterpret_cast<float*>(take((size_t)gt_pack_floats(plan->dims[k], plan->dims[k + 1]) * 4));
    }
  }
  if ((size_t)(base - static_cast<char*>(device_buffer)) > bytes) { delete h; return MOLANN_ERR_WORKSPACE; }
  cudaError_t ce = cudaSuccess;
  if (h->n_pos) ce = cudaMemcpyAsync(h->pos_atom, pos_atom.data(), pos_atom.size() * 4, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess && h->n_inv_ent)
    ce = cudaMemcpyAsync(h->inv_ent, inv_ent.data(), inv_ent.size() * 4, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(h->colmap, colmap.data(), colmap.size() * 4, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);          // the host vectors go out of scope below
  if (ce != cudaSuccess) { delete h; return check_cuda(ce); }
  s = wide_pack_weights(h, plan, st);
  if (s) { delete h; return s; }
  *out = h;
  return MOLANN_OK;
}

int molann_b200_prepared_refresh(MolannPrepared* prepared, const MolannPlan* plan, void* stream) {
  if (!prepared || !plan) return MOLANN_ERR_NULL;
  int s = validate_full(plan);
  if (s) return s;
  if (!prepared_matches(prepared, plan)) return MOLANN_ERR_PLAN;
  return wide_pack_weights(prepared, plan, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------------
// value-and-gradient of a prepared wide plan: forward = the fused wide kernel, which also leaves the hidden
// activations behind (1.5 KB per C3 frame); backward = the layered contractions on the packed operands + the block
// preprocess backward.  The layered route (general_backward) recomputed the forward with four more launches per chunk
// and wrote / re-read the [L, d] feature matrix.  Chunks are whole waves of 148 x 128 frames, as many as ~2 GB of
// workspace hold: the fused kernel wants several tiles per CTA (pipeline fill = one tile).
// ---------------------------------------------------------------------------------------------------------------
long long wide_vg_chunk(const MolannPrepared* h, long long L, int sm_count) {
  long long md = h->d_feat;
  for (int k = 0; k <= h->n_layers; ++k) md = h->dims[k] > md ? h->dims[k] : md;
  const long long per_frame = 4LL * (h->n1 + h->n2 + 2 * md);
  const long long wave = (long long)sm_count * FW_M;
  long long ch = (2LL << 30) / per_frame / wave * wave;
  if (ch < wave) ch = wave;
  ch = env_int("MOLANN_B200_WIDE_VG_CHUNK", (int)ch);
  if (ch < FW_M) ch = FW_M;
  return L < ch ? L : ch;
}
size_t wide_vg_ws_bytes(const MolannPrepared* h, long long L, int sm_count) {
  const long long ch = wide_vg_chunk(h, L, sm_count);
  long long md = h->d_feat;
  for (int k = 0; k <= h->n_layers; ++k) md = h->dims[k] > md ? h->dims[k] : md;
  const long long ntiles = (ch + FW_M - 1) / FW_M;
  const long long grid = sm_count < ntiles ? sm_count : ntiles;
  const long long slot_floats = (long long)FW_SUB * fw_row_floats(h->nkc1, h->ku);
  size_t b = align256((size_t)grid * FW_MAX_SLOTS * slot_floats * 4 + 128);
  b += align256((size_t)ch * h->n1 * 4) + align256((size_t)ch * (h->n2 > 0 ? h->n2 : 1) * 4);
  b += 2 * align256((size_t)ch * md * 4);
  return b;
}
int wide_value_and_grad(const MolannPrepared* h, const MolannPlan* p, const float* x, const float* gy, long long L,
                        float* y, float* gx, void* ws, size_t ws_bytes, const DeviceInfo& dev, cudaStream_t st) {
  if (!ws || ws_bytes < wide_vg_ws_bytes(h, L, dev.sm_count) || (reinterpret_cast<uintptr_t>(ws) & 15u))
    return MOLANN_ERR_WORKSPACE;
  const DevPlan dp = to_dev(p);
  const long long ch = wide_vg_chunk(h, L, dev.sm_count);
  const int nl = p->n_layers, kout = p->dims[nl];
  long long md = p->d_feat;
  for (int k = 0; k <= nl; ++k) md = p->dims[k] > md ? p->dims[k] : md;
  char* base = static_cast<char*>(ws);
  {
    const long long ntiles = (ch + FW_M - 1) / FW_M;
    const long long grid = dev.sm_count < ntiles ? dev.sm_count : ntiles;
    const long long slot_floats = (long long)FW_SUB * fw_row_floats(h->nkc1, h->ku);
    base += align256((size_t)grid * FW_MAX_SLOTS * slot_floats * 4 + 128);
  }
  float* hact[MOLANN_MAX_LAYERS] = {nullptr};          // hact[k] = activation after layer k (k = 1 .. nl - 1)
  hact[1] = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * h->n1 * 4);
  float* h2buf = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * (h->n2 > 0 ? h->n2 : 1) * 4);
  if (nl == 3) hact[2] = h2buf;
  float* pp[2];
  pp[0] = reinterpret_cast<float*>(base);
  base += align256((size_t)ch * md * 4);
  pp[1] = reinterpret_cast<float*>(base);
  for (long long c0 = 0; c0 < L; c0 += ch) {
    const long long Lc = (L - c0 < ch) ? (L - c0) : ch;
    const float* xc = x + c0 * 3 * p->n_inp;
    WideChoice wc = choose_wide(h, p, Lc, dev);
    if (!wc.ok) return MOLANN_ERR_PLAN;
    wc.P.scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 127u) & ~(uintptr_t)127u);
    int s = launch_wide(wc, h->ku, p->act_id, xc, y + c0 * kout, Lc, st, hact[1], nl == 3 ? hact[2] : nullptr);
    if (s) return s;
    const float* gz = gy + c0 * kout;
    for (int k = nl - 1; k >= 0; --k) {
      float* gprev = pp[k & 1];
      s = launch_linear_backward_input(gz, p->W[k], k > 0 ? hact[k] : nullptr, gprev, Lc, p->dims[k], p->dims[k + 1],
                                       p->act_id, st, h->gt_bwd[k], &dev, h->gt_bwd[k]);   // (pack != NULL selects the tensor-core GEMM; nothing is packed here)
      if (s) return s;
      gz = gprev;
    }
    s = launch_preprocess_backward(p, dp, xc, gz, gx + c0 * 3 * p->n_inp, Lc, dev, st);
    if (s) return s;
  }
  return MOLANN_OK;
}

size_t molann_b200_prepared_workspace_bytes(const MolannPrepared* prepared, int64_t L) {
  if (!prepared || prepared->magic != kPreparedMagic || L <= 0) return 0;
  DeviceInfo dev = device_info();
  if (!dev.ok) dev.sm_count = 148;
  const long long ntiles = (L + FW_M - 1) / FW_M;
  const long long grid = dev.sm_count < ntiles ? dev.sm_count : ntiles;
  const long long slot_floats = (long long)FW_SUB * fw_row_floats(prepared->nkc1, prepared->ku);
  const size_t wide = (size_t)grid * FW_MAX_SLOTS * slot_floats * 4 + 128;    // + alignment of the rows to L2 lines
  // the layered kernels (value-and-gradient; forward when the wide kernel does not fit) share the same workspace
  MolannPlan shape;
  std::memset(&shape, 0, sizeof(shape));
  shape.d_feat = prepared->d_feat;
  shape.n_layers = prepared->n_layers;
  for (int k = 0; k <= prepared->n_layers; ++k) shape.dims[k] = prepared->dims[k];
  const size_t layered = general_ws_bytes(&shape, L, true);
  const size_t vg = wide_vg_ws_bytes(prepared, L, dev.sm_count);
  size_t m = wide > layered ? wide : layered;
  return m > vg ? m : vg;
}

int molann_b200_forward_prepared(const MolannPrepared* prepared, const MolannPlan* plan, const float* x, int64_t L,
                                 float* y, void* workspace, size_t workspace_bytes, void* stream) {
  int s = validate_full(plan);
  if (s) return s;
  if (!prepared_matches(prepared, plan)) return MOLANN_ERR_PLAN;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !y) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(y)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  WideChoice ch = choose_wide(prepared, plan, (long long)L, dev);
  if (!ch.ok || env_int("MOLANN_B200_WIDE", -1) == 0)      // e.g. 60 KB frames: layered kernels on the packed operands
    return general_forward(plan, x, L, y, workspace, workspace_bytes, dev, static_cast<cudaStream_t>(stream), prepared);
  const size_t need = (size_t)ch.grid * (size_t)ch.P.cta_floats * 4 + 128;
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15u)) return MOLANN_ERR_WORKSPACE;
  // rows start on 128-byte L2 lines: a converter thread discards the lines of its row once it has read them
  ch.P.scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 127u) & ~(uintptr_t)127u);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return launch_wide(ch, prepared->ku, plan->act_id, x, y, (long long)L, st, nullptr, nullptr);
}

int molann_b200_value_and_grad_prepared(const MolannPrepared* prepared, const MolannPlan* plan, const float* x,
                                        const float* gy, int64_t L, float* y, float* gx, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  int s = validate_full(plan);
  if (s) return s;
  if (!prepared_matches(prepared, plan)) return MOLANN_ERR_PLAN;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !gy || !y || !gx) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(gy) || misaligned4(y) || misaligned4(gx)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  // forward on the fused wide kernel (which keeps the hidden activations), backward on the layered kernels; without
  // the wide kernel: ONE pass of the layered kernels (the backward's forward recompute also writes y)
  if (env_int("MOLANN_B200_WIDE", -1) != 0 && env_int("MOLANN_B200_WIDE_VG", 1) != 0 && wide_can_store_h(plan->act_id) &&
      choose_wide(prepared, plan, (long long)L, dev).ok)
    return wide_value_and_grad(prepared, plan, x, gy, (long long)L, y, gx, workspace, workspace_bytes, dev,
                               static_cast<cudaStream_t>(stream));
  return general_backward(plan, x, gy, L, gx, nullptr, nullptr, workspace, workspace_bytes, dev,
                          static_cast<cudaStream_t>(stream), y, prepared);
}

void molann_b200_prepared_destroy(MolannPrepared* prepared) {
  if (prepared && prepared->magic == kPreparedMagic) {
    prepared->magic = 0;
    delete prepared;
  }
}

int molann_b200_decode_frames_i16(const int16_t* q, int64_t n_values, const float* origin_host, float resolution,
                                  float* x, void* stream) {
  if (n_values < 0 || n_values % 3 != 0) return MOLANN_ERR_PLAN;
  if (n_values == 0) return MOLANN_OK;
  if (!q || !x || !origin_host) return MOLANN_ERR_NULL;
  if ((reinterpret_cast<uintptr_t>(q) & 1u) || misaligned4(x)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  const int vec = ((reinterpret_cast<uintptr_t>(q) & 15u) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
  long long blocks = (n_values / 8 + 255) / 256;
  const long long cap = (long long)dev.sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  decode_frames_i16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q, (long long)n_values, origin_host[0], origin_host[1], origin_host[2], resolution, x, vec);
  return post_launch();
}

size_t molann_b200_jacobian_workspace_bytes(const MolannPlan* plan, int64_t L) {
  if (validate_full(plan) != MOLANN_OK || plan->n_layers < 1 || L <= 0) return 0;
  return general_ws_bytes(plan, L, true) + align256((size_t)L * plan->dims[plan->n_layers] * 4);
}

int molann_b200_value_and_jacobian(const MolannPlan* plan, const float* x, int64_t L, float* y, float* jac,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  int s = validate_full(plan);
  if (s) return s;
  if (plan->n_layers < 1) return MOLANN_ERR_UNSUPPORTED;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !y || !jac) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(y) || misaligned4(jac)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kout = plan->dims[plan->n_layers];
  if (env_int("MOLANN_B200_JAC_FUSED", 1) != 0) {
    const TcVgChoice vg = choose_tc_vg(plan, dev, 1);
    if (vg.ok && vg.tiles == 1) {                  // ONE launch: every plane while the tile is on chip
      const DevPlan dp = to_dev(plan);
      switch (plan->act_id) {
        case MOLANN_ACT_TANH: return launch_tc_jac<ACT_TANH>(dp, vg.vl, x, y, jac, (long long)L, dev, st);
        case MOLANN_ACT_RELU: return launch_tc_jac<ACT_RELU>(dp, vg.vl, x, y, jac, (long long)L, dev, st);
        case MOLANN_ACT_SIGMOID: return launch_tc_jac<ACT_SIGMOID>(dp, vg.vl, x, y, jac, (long long)L, dev, st);
        default: return launch_tc_jac<ACT_IDENTITY>(dp, vg.vl, x, y, jac, (long long)L, dev, st);
      }
    }
  }
  // other kernel families: one value-and-gradient pass per output with a one-hot cotangent
  const size_t cot_bytes = align256((size_t)L * kout * 4);
  if (!workspace || workspace_bytes < molann_b200_jacobian_workspace_bytes(plan, L)) return MOLANN_ERR_WORKSPACE;
  float* cot = static_cast<float*>(workspace);
  char* rest = static_cast<char*>(workspace) + cot_bytes;
  for (int o = 0; o < kout; ++o) {
    unsigned blocks = (unsigned)(((long long)L * kout + 255) / 256);
    if (blocks > 1184u) blocks = 1184u;
    onehot_rows_kernel<<<blocks, 256, 0, st>>>(cot, (long long)L, kout, o);
    s = post_launch();
    if (s) return s;
    s = molann_b200_value_and_grad(plan, x, cot, L, y, jac + (size_t)o * (size_t)L * 3 * plan->n_inp, rest,
                                   workspace_bytes - cot_bytes, stream);
    if (s) return s;
  }
  return MOLANN_OK;
}

int molann_b200_preprocess_forward(const MolannPlan* plan, const float* x, int64_t L, float* feat, void* stream) {
  int s = validate_features(plan);
  if (s) return s;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !feat) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(feat)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  return launch_preprocess_forward(plan, to_dev(plan), x, feat, (long long)L, dev, static_cast<cudaStream_t>(stream));
}

int molann_b200_preprocess_backward(const MolannPlan* plan, const float* x, const float* gfeat, int64_t L, float* gx,
                                    void* stream) {
  int s = validate_features(plan);
  if (s) return s;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !gfeat || !gx) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(gfeat) || misaligned4(gx)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  return launch_preprocess_backward(plan, to_dev(plan), x, gfeat, gx, (long long)L, dev,
                                    static_cast<cudaStream_t>(stream));
}

int molann_b200_align_forward(const MolannPlan* plan, const float* x, int64_t L, float* out, void* stream) {
  int s = validate_geometry(plan);
  if (s) return s;
  if (plan->n_align < 1) return MOLANN_ERR_PLAN;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !out) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(out)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  DevPlan dp = to_dev(plan);
  dp.n_entries = 0;                              // only the geometry fields of the plan are meaningful here
  const TileChoice tc = choose_align_tile(plan, false, dev);
  if (tc.ok) {
    s = check_cuda(cudaFuncSetAttribute(align_forward_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        tc.lay.total));
    if (s) return s;
    align_forward_tile_kernel<<<tile_grid(L, tc.lay.total, dev), ST_F, tc.lay.total,
                                static_cast<cudaStream_t>(stream)>>>(dp, tc.lay, x, out, L);
    return post_launch();
  }
  align_forward_warp_kernel<<<warp_grid(L, dev), WARPS_PER_CTA * 32, 0, static_cast<cudaStream_t>(stream)>>>(dp, x,
                                                                                                             out, L);
  return post_launch();
}

int molann_b200_align_backward(const MolannPlan* plan, const float* x, const float* gout, int64_t L, float* gx,
                               void* stream) {
  int s = validate_geometry(plan);
  if (s) return s;
  if (plan->n_align < 1) return MOLANN_ERR_PLAN;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (L == 0) return MOLANN_OK;
  if (!x || !gout || !gx) return MOLANN_ERR_NULL;
  if (misaligned4(x) || misaligned4(gout) || misaligned4(gx)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  DevPlan dp = to_dev(plan);
  dp.n_entries = 0;
  const TileChoice tc = choose_align_tile(plan, true, dev);
  if (tc.ok) {
    s = check_cuda(cudaFuncSetAttribute(align_backward_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        tc.lay.total));
    if (s) return s;
    align_backward_tile_kernel<<<tile_grid(L, tc.lay.total, dev), ST_F, tc.lay.total,
                                 static_cast<cudaStream_t>(stream)>>>(dp, tc.lay, x, gout, gx, L);
    return post_launch();
  }
  align_backward_warp_kernel<<<warp_grid(L, dev), WARPS_PER_CTA * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      dp, x, gout, gx, L);
  return post_launch();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// autoencoder training step (fused_train.cuh)
// ---------------------------------------------------------------------------------------------
namespace {

struct TrainChoice {
  bool ok = false;
  TrainNet net;
  TrainLayout lay;
};

int validate_decoder(const MolannPlan* enc, const MolannDecoder* dec) {
  int s = validate_full(enc);
  if (s) return s;
  if (!dec) return MOLANN_ERR_NULL;
  if (enc->n_layers < 1 || dec->n_layers < 1 || dec->n_layers > MOLANN_MAX_LAYERS) return MOLANN_ERR_PLAN;
  if (dec->act_id < 0 || dec->act_id > MOLANN_ACT_IDENTITY) return MOLANN_ERR_PLAN;
  if (dec->dims[0] != enc->dims[enc->n_layers] || dec->dims[dec->n_layers] != enc->d_feat) return MOLANN_ERR_PLAN;
  for (int k = 0; k <= dec->n_layers; ++k)
    if (dec->dims[k] <= 0) return MOLANN_ERR_PLAN;
  for (int k = 0; k < dec->n_layers; ++k)
    if (!dec->W[k] || !dec->b[k]) return MOLANN_ERR_NULL;
  return MOLANN_OK;
}

TrainChoice choose_train(const MolannPlan* enc, const MolannDecoder* dec, const DeviceInfo& dev) {
  TrainChoice tc;
  std::memset(&tc.net, 0, sizeof(tc.net));
  std::memset(&tc.lay, 0, sizeof(tc.lay));
  TrainNet& n = tc.net;
  n.ne = enc->n_layers;
  n.nl = enc->n_layers + dec->n_layers;
  n.act_enc = enc->act_id;
  n.act_dec = dec->act_id;
  long long P = 0;
  for (int l = 0; l < n.nl; ++l) {
    const bool e = l < n.ne;
    const int k = e ? l : l - n.ne;
    n.c[l] = e ? enc->dims[k] : dec->dims[k];
    n.c[l + 1] = e ? enc->dims[k + 1] : dec->dims[k + 1];
    n.W[l] = e ? enc->W[k] : dec->W[k];
    n.b[l] = e ? enc->b[k] : dec->b[k];
    n.gw[l] = (int)P;
    P += (long long)n.c[l] * n.c[l + 1];
    n.gb[l] = (int)P;
    P += n.c[l + 1];
    if (P > (1ll << 28)) return tc;
  }
  n.P = (int)P;
  TrainLayout& lay = tc.lay;
  long long off = 0;
  lay.act_lo = 0;
  for (int l = 0; l <= n.nl; ++l) {
    lay.a_off[l] = (int)off;
    off += (long long)round_up(n.c[l], 4) * TR_FS * 4;
  }
  lay.act_bytes = (int)off;
  const long long tile_bytes = (long long)TR_F * 3 * enc->n_inp * 4;
  const long long over = (off - tile_bytes) / 128 * 128;       // overlay on the tail (dead rows during the geometry)
  if (off - tile_bytes >= 0 && over >= lay.a_off[1]) {
    lay.xs_off = (int)over;
  } else {
    off = (off + 127) / 128 * 128;
    lay.xs_off = (int)off;
    off += tile_bytes;
  }
  for (int l = 0; l < n.nl; ++l) {
    lay.ldk[l] = round_up(n.c[l], 8);
    off = (off + 15) / 16 * 16;
    lay.w_off[l] = (int)off;
    off += (long long)round_up(n.c[l + 1], 8) * lay.ldk[l] * 4;
    lay.b_off[l] = (int)off;
    off += (long long)round_up(n.c[l + 1], 8) * 4;
    if (off > (1ll << 24)) return tc;
  }
  // work shapes (fused_train.cuh): keep all eight warps busy in every phase
  for (int l = 0; l < n.nl; ++l) {
    const int K = n.c[l], N = n.c[l + 1];
    lay.fw[l] = N <= 4 ? TR_NARROW : (N > 56 ? TR_WIDE : TR_HALF);
    lay.bw[l] = K <= 4 ? TR_NARROW : (K > 56 ? TR_WIDE : TR_HALF);
    lay.dw[l] = TR_DW22;
    if (N <= 4 || K <= 4) {
      lay.dw[l] = TR_DWTHIN;
    } else {
      const int ro[3] = {4, 4, 2}, ri[3] = {4, 2, 4}, mode[3] = {TR_DW44, TR_DW42, TR_DW24};
      for (int c = 0; c < 3; ++c) {
        const int TO = (N + ro[c] - 1) / ro[c], TI = (K + ri[c] - 1) / ri[c];
        if (((TO + 7) / 8) * ((TI + 3) / 4) >= TR_NT / 32) {
          lay.dw[l] = mode[c];
          break;
        }
      }
      // large tiles over frame quarters when they fill the lanes (6 .. 8 column groups) and all eight warps
      if (env_int("MOLANN_B200_TRAIN_BIG", 1) != 0) {
        const int bo[4] = {8, 8, 4, 4}, bi[4] = {8, 4, 8, 4}, bm[4] = {TR_DWB88, TR_DWB84, TR_DWB48, TR_DWB44};
        for (int c = 0; c < 4; ++c) {
          const int TO = (N + bo[c] - 1) / bo[c], TI = (K + bi[c] - 1) / bi[c];
          if (TI >= 6 && TI <= 8 && TO >= 8 && TO % 8 == 0) {
            lay.dw[l] = bm[c];
            break;
          }
        }
      }
    }
  }
  off = (off + 15) / 16 * 16;
  lay.aidx_off = (int)off; off += (long long)round_up(enc->n_align > 0 ? enc->n_align : 1, 4) * 4;
  lay.ref_off = (int)off; off += (long long)round_up(3 * (enc->n_align > 0 ? enc->n_align : 1), 4) * 4;
  lay.ent_off = (int)off; off += (long long)round_up(ENTRY_INTS * enc->n_entries, 4) * 4;
  lay.mbar_off = (int)off; off += 16;
  lay.red_off = (int)off; off += (TR_NT / 32) * 4;
  off = (off + 15) / 16 * 16;
  lay.total_bytes = (int)off;
  tc.ok = off <= dev.max_smem_optin;
  return tc;
}

}  // namespace

extern "C" {

int molann_b200_train_eligible(const MolannPlan* encoder, const MolannDecoder* decoder) {
  if (validate_decoder(encoder, decoder) != MOLANN_OK) return 0;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return 0;
  return choose_train(encoder, decoder, dev).ok ? 1 : 0;
}

size_t molann_b200_train_param_count(const MolannPlan* encoder, const MolannDecoder* decoder) {
  if (validate_decoder(encoder, decoder) != MOLANN_OK) return 0;
  size_t P = 0;
  for (int k = 0; k < encoder->n_layers; ++k) P += (size_t)(encoder->dims[k] + 1) * encoder->dims[k + 1];
  for (int k = 0; k < decoder->n_layers; ++k) P += (size_t)(decoder->dims[k] + 1) * decoder->dims[k + 1];
  return P;
}

size_t molann_b200_train_workspace_bytes(const MolannPlan* encoder, const MolannDecoder* decoder) {
  const size_t P = molann_b200_train_param_count(encoder, decoder);
  const DeviceInfo dev = device_info();
  if (!P || !dev.ok) return 0;
  return align256((size_t)dev.sm_count * (P + 1) * 4);
}

int molann_b200_train_loss_and_grads(const MolannPlan* encoder, const MolannDecoder* decoder, const float* x, int64_t L,
                                     float loss_scale, float* flat, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  int s = validate_decoder(encoder, decoder);
  if (s) return s;
  if (L < 0) return MOLANN_ERR_PLAN;
  if (!flat) return MOLANN_ERR_NULL;
  if (misaligned4(flat)) return MOLANN_ERR_ALIGNMENT;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TrainChoice tc = choose_train(encoder, decoder, dev);
  if (!tc.ok) return MOLANN_ERR_UNSUPPORTED;
  const size_t n_flat = (size_t)tc.net.P + 1;
  if (L == 0) return check_cuda(cudaMemsetAsync(flat, 0, n_flat * 4, st));
  if (!x) return MOLANN_ERR_NULL;
  if (misaligned4(x)) return MOLANN_ERR_ALIGNMENT;
  if (!workspace || workspace_bytes < molann_b200_train_workspace_bytes(encoder, decoder)) return MOLANN_ERR_WORKSPACE;
  if (misaligned4(workspace)) return MOLANN_ERR_ALIGNMENT;
  const long long ntiles = (L + TR_F - 1) / TR_F;
  const int grid = (int)(ntiles < dev.sm_count ? ntiles : dev.sm_count);     // every CTA owns at least one tile
  s = check_cuda(cudaFuncSetAttribute(fused_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      tc.lay.total_bytes));
  if (s) return s;
  const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && env_int("MOLANN_B200_TMA", 1) != 0) ? 1 : 0;
  float* planes = static_cast<float*>(workspace);
  fused_train_kernel<<<grid, TR_NT, tc.lay.total_bytes, st>>>(to_dev(encoder), tc.net, tc.lay, x, (long long)L,
                                                               loss_scale, planes, use_tma);
  s = post_launch();
  if (s) return s;
  train_reduce_kernel<<<(unsigned)((n_flat + TR_RED_X - 1) / TR_RED_X), dim3(TR_RED_X, TR_RED_Y), 0, st>>>(planes, grid, (int)n_flat,
                                                                                                          flat);
  return post_launch();
}

int molann_b200_sgd_apply(float* const* params, const int64_t* numel, int32_t n_params, const float* flat, float lr,
                          void* stream) {
  if (!params || !numel || !flat) return MOLANN_ERR_NULL;
  if (n_params < 0 || n_params > 2 * TR_MAXL) return MOLANN_ERR_PLAN;
  if (n_params == 0) return MOLANN_OK;
  SgdTable tab;
  std::memset(&tab, 0, sizeof(tab));
  long long total = 0;
  for (int i = 0; i < n_params; ++i) {
    if (!params[i]) return MOLANN_ERR_NULL;
    if (numel[i] < 0 || misaligned4(params[i])) return numel[i] < 0 ? MOLANN_ERR_PLAN : MOLANN_ERR_ALIGNMENT;
    total += numel[i];
    if (total > (1ll << 30)) return MOLANN_ERR_PLAN;
    tab.ptr[i] = params[i];
    tab.end[i] = (int)total;
  }
  tab.n = n_params;
  if (total == 0) return MOLANN_OK;
  if (!device_info().ok) return MOLANN_ERR_CUDA;
  train_sgd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(tab, flat, lr,
                                                                                                   (int)total);
  return post_launch();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// one-shot allreduce over peer memory + SGD (fused_train.cuh)
// ---------------------------------------------------------------------------------------------
extern "C" {

size_t molann_b200_allreduce_buffer_bytes(int64_t n, int32_t world) {
  if (n <= 0 || world < 1 || world > TR_MAX_PEERS) return 0;
  return align256((size_t)2 * (size_t)n * 4) + 256;              // two slots, then the flag row (world uint32)
}

int molann_b200_allreduce_sgd(const float* flat_local, float* flat_global, int64_t n, void* const* peer_buffers,
                              int32_t rank, int32_t world, uint32_t* state, float* const* params,
                              const int64_t* numel, int32_t n_params, float lr, void* stream) {
  if (!flat_local || !flat_global || !peer_buffers || !state) return MOLANN_ERR_NULL;
  if (n <= 0 || n > (1ll << 28) || world < 1 || world > TR_MAX_PEERS || rank < 0 || rank >= world) return MOLANN_ERR_PLAN;
  if (n_params < 0 || n_params > 2 * TR_MAXL || (n_params > 0 && (!params || !numel))) {
    return n_params < 0 || n_params > 2 * TR_MAXL ? MOLANN_ERR_PLAN : MOLANN_ERR_NULL;
  }
  if (misaligned4(flat_local) || misaligned4(flat_global) || misaligned4(state)) return MOLANN_ERR_ALIGNMENT;
  PeerTable pt;
  std::memset(&pt, 0, sizeof(pt));
  const size_t flag_off = align256((size_t)2 * (size_t)n * 4);
  for (int r = 0; r < world; ++r) {
    if (!peer_buffers[r]) return MOLANN_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(peer_buffers[r]) & 15u) return MOLANN_ERR_ALIGNMENT;
    pt.buf[r] = static_cast<float*>(peer_buffers[r]);
    pt.flag[r] = reinterpret_cast<unsigned*>(static_cast<char*>(peer_buffers[r]) + flag_off);
  }
  pt.rank = rank;
  pt.world = world;
  SgdTable tab;
  std::memset(&tab, 0, sizeof(tab));
  long long total = 0;
  for (int i = 0; i < n_params; ++i) {
    if (!params[i]) return MOLANN_ERR_NULL;
    if (numel[i] < 0) return MOLANN_ERR_PLAN;
    if (misaligned4(params[i])) return MOLANN_ERR_ALIGNMENT;
    total += numel[i];
    if (total > n) return MOLANN_ERR_PLAN;
    tab.ptr[i] = params[i];
    tab.end[i] = (int)total;
  }
  tab.n = n_params > 0 ? n_params : 1;
  const DeviceInfo dev = device_info();
  if (!dev.ok) return MOLANN_ERR_CUDA;
  long long blocks = (n + 255) / 256;
  if (blocks > dev.sm_count) blocks = dev.sm_count;              // co-resident: CTAs wait for each other
  train_allreduce_sgd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pt, tab, flat_local, flat_global, (int)n, (int)total, n_params > 0 ? lr : 0.f, state);
  return post_launch();
}

}  // extern "C"
