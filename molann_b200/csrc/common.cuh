// common.cuh -- plan mirror passed to kernels by value, shared-memory layout of the fused kernels,
// and the sm_100a async-copy (TMA bulk) / mbarrier primitives used for tile staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/molann_b200.h"

namespace molann {

struct DevPlan {
  int n_inp, n_align, n_entries, d_feat, use_angle, n_layers, act;
  const int* align_idx;
  const float* ref_x;
  const int* entries;
  int dims[MOLANN_MAX_LAYERS + 1];
  const float* W[MOLANN_MAX_LAYERS];
  const float* b[MOLANN_MAX_LAYERS];
};

// Byte offsets into dynamic shared memory for the fused small-system kernels (host-computed).
struct SmallLayout {
  int xs_off;                       // staged coordinate tile  [F][3n]
  int gxs_off;                      // gradient tile           [F][3n]   (backward only)
  int buf_off[MOLANN_MAX_LAYERS];   // activation buffers      [rows][F]
  int gys_off;                      // output cotangent tile   [kout_pad][F] (backward only)
  int w_off[MOLANN_MAX_LAYERS];     // natural weights  Wn[o][i]  (backward only), ld = ldw
  int wt_off[MOLANN_MAX_LAYERS];    // transposed weights Wt[i][o] (forward / recompute), ld = ldwt
  int b_off[MOLANN_MAX_LAYERS];     // biases (padded)
  int ldw[MOLANN_MAX_LAYERS];
  int ldwt[MOLANN_MAX_LAYERS];
  int tm_fwd[MOLANN_MAX_LAYERS];    // frames per micro-tile, forward layer k
  int tm_bwd[MOLANN_MAX_LAYERS];    // frames per micro-tile, backward layer k
  int aidx_off, ref_off, ent_off, mbar_off;
  int alias_xs;                     // forward: activation buffer 1 overlays the coordinate tile
  int total_bytes;
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP).  A whole frame tile is one contiguous
// byte range of x, so one bulk copy per tile stages it without touching registers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared, completion signalled on the mbarrier (bytes % 16 == 0, both addresses 16B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global (bulk-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace molann
