// torch_shim.cpp -- thin dispatcher-registered custom ops over the C ABI (include/molann_b200.h).
//
//   molann_b200::align       <- AlignmentLayer.forward      (reference molann/ann.py:157-199)
//   molann_b200::preprocess  <- PreprocessingANN.forward    (:553-565; FeatureLayer.forward :454-474 when
//                                                            align_idx is empty)
//   molann_b200::molann      <- MolANN.forward              (:620-624) with a create_sequential_nn MLP
//
// Autograd is implemented in C++ (torch::autograd::Function) so that TorchScript archives loaded from
// libtorch-only consumers (MD-engine plugins) still get d/dx and parameter gradients.  PyTorch is
// plumbing here: tensors for device memory, the current stream, and the dispatcher; all arithmetic is
// in the sm_100a kernels behind the C ABI.  There is no CPU implementation: CPU tensors raise.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <ATen/ATen.h>
#include <torch/csrc/autograd/custom_function.h>

#include <nvtx3/nvToolsExt.h>

#include <cstring>
#include <list>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/molann_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check_status(int status, const char* what) {
  if (status == MOLANN_OK) return;
  if (status == MOLANN_ERR_CUDA) {
    TORCH_CHECK(false, "molann_b200: ", what, " failed: ", molann_b200_strerror(status), " (cudaError ",
                molann_b200_last_cuda_error(), ": ", molann_b200_cuda_error_string(molann_b200_last_cuda_error()),
                ")");
  }
  TORCH_CHECK(false, "molann_b200: ", what, " failed: ", molann_b200_strerror(status));
}

// NVTX range around every op (tracing); a no-op unless a profiler is attached
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// The backward kernels return plain tensors: a second differentiation through them (create_graph=True, e.g. a
// force-matching or gradient-norm loss) would silently treat d/dx as a constant.  The autograd engine enables grad
// mode inside backward only for create_graph=True, so that is the condition to refuse.
void refuse_double_backward(const char* op) {
  TORCH_CHECK(!at::GradMode::is_enabled(), "molann_b200::", op,
              ": double backward (torch.autograd.grad(..., create_graph=True)) is not supported -- the coordinate "
              "gradient comes from a closed-form kernel and carries no autograd history");
}

void check_cotangent(const Tensor& g, const Tensor& x, const char* op) {
  TORCH_CHECK(g.defined(), "molann_b200::", op, ": undefined output gradient");
  TORCH_CHECK(g.device() == x.device(), "molann_b200::", op, ": the output gradient lives on ", g.device(),
              " but the input on ", x.device());
  TORCH_CHECK(g.scalar_type() == at::kFloat, "molann_b200::", op, ": the output gradient must be float32, got ",
              g.scalar_type());
}

void check_x(const Tensor& x, const char* op) {
  TORCH_CHECK(x.is_cuda(), "molann_b200::", op,
              ": input must be a CUDA tensor (this build has no CPU implementation); got device ", x.device());
  TORCH_CHECK(x.scalar_type() == at::kFloat, "molann_b200::", op, ": input must be float32, got ", x.scalar_type());
  TORCH_CHECK(x.dim() == 3 && x.size(2) == 3, "molann_b200::", op, ": input must have shape [L, n_inp, 3]");
  TORCH_CHECK(x.is_contiguous(), "molann_b200::", op, ": input must be contiguous");
}

void check_const(const Tensor& t, const Tensor& x, at::ScalarType dt, const char* name) {
  if (t.numel() == 0) return;
  TORCH_CHECK(t.device() == x.device(), "molann_b200: ", name, " must live on the input's device (", x.device(),
              "), got ", t.device(), " -- call module.to(device)");
  TORCH_CHECK(t.scalar_type() == dt, "molann_b200: ", name, " has wrong dtype ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous(), "molann_b200: ", name, " must be contiguous");
}

struct PlanHolder {
  MolannPlan plan;
  std::vector<Tensor> keep;      // contiguous views kept alive for the duration of the call
};

void fill_geometry(PlanHolder& h, const Tensor& x, const Tensor& align_idx, const Tensor& ref_x) {
  std::memset(&h.plan, 0, sizeof(MolannPlan));
  check_const(align_idx, x, at::kInt, "align_idx");
  check_const(ref_x, x, at::kFloat, "ref_x");
  h.plan.n_inp = static_cast<int32_t>(x.size(1));
  h.plan.n_align = static_cast<int32_t>(align_idx.numel());
  TORCH_CHECK(ref_x.numel() == 3 * align_idx.numel(), "molann_b200: ref_x must be [n_align, 3]");
  h.plan.align_idx = h.plan.n_align ? align_idx.data_ptr<int32_t>() : nullptr;
  h.plan.ref_x = h.plan.n_align ? ref_x.data_ptr<float>() : nullptr;
}

void fill_features(PlanHolder& h, const Tensor& x, const Tensor& entries, int64_t d_feat, bool use_angle_value) {
  check_const(entries, x, at::kInt, "feature program");
  TORCH_CHECK(entries.dim() == 2 && entries.size(1) == MOLANN_ENTRY_INTS && entries.size(0) > 0,
              "molann_b200: feature program must be a non-empty int32 [n_entries, ", MOLANN_ENTRY_INTS, "] tensor");
  h.plan.n_entries = static_cast<int32_t>(entries.size(0));
  h.plan.entries = entries.data_ptr<int32_t>();
  h.plan.d_feat = static_cast<int32_t>(d_feat);
  h.plan.use_angle_value = use_angle_value ? 1 : 0;
}

void fill_mlp(PlanHolder& h, const Tensor& x, at::TensorList params, int64_t act) {
  TORCH_CHECK(params.size() >= 2 && params.size() % 2 == 0, "molann_b200: params must be [W1, b1, W2, b2, ...]");
  const int nl = static_cast<int>(params.size() / 2);
  TORCH_CHECK(nl <= MOLANN_MAX_LAYERS, "molann_b200: at most ", MOLANN_MAX_LAYERS, " linear layers are supported");
  h.plan.n_layers = nl;
  h.plan.act_id = static_cast<int32_t>(act);
  h.plan.dims[0] = h.plan.d_feat;
  for (int k = 0; k < nl; ++k) {
    const Tensor& W = params[2 * k];
    const Tensor& b = params[2 * k + 1];
    TORCH_CHECK(W.dim() == 2 && b.dim() == 1 && W.size(0) == b.size(0), "molann_b200: bad Linear shapes at layer ",
                k + 1);
    TORCH_CHECK(W.size(1) == h.plan.dims[k], "molann_b200: layer ", k + 1, " expects ", W.size(1),
                " inputs but receives ", h.plan.dims[k]);
    TORCH_CHECK(W.device() == x.device() && b.device() == x.device(),
                "molann_b200: MLP parameters must live on the input's device -- call module.to(device)");
    TORCH_CHECK(W.scalar_type() == at::kFloat && b.scalar_type() == at::kFloat,
                "molann_b200: MLP parameters must be float32");
    h.keep.push_back(W.contiguous());
    h.keep.push_back(b.contiguous());
    h.plan.W[k] = h.keep[2 * k].data_ptr<float>();
    h.plan.b[k] = h.keep[2 * k + 1].data_ptr<float>();
    h.plan.dims[k + 1] = static_cast<int32_t>(W.size(0));
  }
}

void* cur_stream() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

// ------------------------------------------------------------------------------------------
// prepared plans (include/molann_b200.h): built on first use of a (feature program, MLP) pair and kept in a small
// LRU; when only the weights changed (training: a new `_version`, or new storage) they are re-packed in place.
// ------------------------------------------------------------------------------------------
struct PreparedEntry {
  int device = -1;
  const void* entries = nullptr;
  const void* align = nullptr;
  const void* ref = nullptr;
  int64_t n_inp = 0, d_feat = 0, act = 0;
  bool use_angle = false;
  std::vector<const void*> param_ptr;
  std::vector<int64_t> param_version;
  std::vector<int64_t> dims;
  Tensor buffer;
  MolannPrepared* handle = nullptr;
  bool unsupported = false;      // the fused wide kernel does not fit this plan on this device: use the layered path
  ~PreparedEntry() { molann_b200_prepared_destroy(handle); }
};
std::mutex g_prepared_mutex;
std::list<std::shared_ptr<PreparedEntry>> g_prepared;          // most recently used first
constexpr size_t kPreparedCacheSize = 16;

std::shared_ptr<PreparedEntry> prepared_for(const PlanHolder& h, const Tensor& x, const Tensor& entries,
                                            const Tensor& align_idx, const Tensor& ref_x, at::TensorList params,
                                            int64_t act, bool use_angle_value) {
  std::lock_guard<std::mutex> lock(g_prepared_mutex);
  const int device = x.get_device();
  std::vector<int64_t> dims(h.plan.dims, h.plan.dims + h.plan.n_layers + 1);
  for (auto it = g_prepared.begin(); it != g_prepared.end(); ++it) {
    PreparedEntry& e = **it;
    if (e.device != device || e.entries != entries.data_ptr() || e.n_inp != h.plan.n_inp || e.d_feat != h.plan.d_feat ||
        e.act != act || e.use_angle != use_angle_value || e.dims != dims ||
        e.align != (align_idx.numel() ? align_idx.data_ptr() : nullptr) ||
        e.ref != (ref_x.numel() ? ref_x.data_ptr() : nullptr))
      continue;
    bool same_weights = e.param_ptr.size() == params.size();
    for (size_t i = 0; same_weights && i < params.size(); ++i)
      same_weights = e.param_ptr[i] == params[i].data_ptr() && e.param_version[i] == (int64_t)params[i]._version();
    if (!same_weights) {
      check_status(molann_b200_prepared_refresh(e.handle, &h.plan, cur_stream()), "prepared_refresh");
      e.param_ptr.clear();
      e.param_version.clear();
      for (const Tensor& p : params) {
        e.param_ptr.push_back(p.data_ptr());
        e.param_version.push_back((int64_t)p._version());
      }
    }
    auto keep = *it;
    g_prepared.erase(it);
    g_prepared.push_front(keep);
    return keep;
  }
  auto e = std::make_shared<PreparedEntry>();
  e->device = device;
  e->entries = entries.data_ptr();
  e->align = align_idx.numel() ? align_idx.data_ptr() : nullptr;
  e->ref = ref_x.numel() ? ref_x.data_ptr() : nullptr;
  e->n_inp = h.plan.n_inp; e->d_feat = h.plan.d_feat; e->act = act; e->use_angle = use_angle_value;
  e->dims = dims;
  for (const Tensor& p : params) {
    e->param_ptr.push_back(p.data_ptr());
    e->param_version.push_back((int64_t)p._version());
  }
  const size_t bytes = molann_b200_prepared_bytes(&h.plan);
  TORCH_CHECK(bytes > 0, "molann_b200: plan is not eligible for the prepared path");
  e->buffer = at::empty({static_cast<int64_t>(bytes)}, x.options().dtype(at::kByte));
  check_status(molann_b200_prepare(&h.plan, e->buffer.data_ptr(), bytes, cur_stream(), &e->handle), "prepare");
  g_prepared.push_front(e);
  while (g_prepared.size() > kPreparedCacheSize) g_prepared.pop_back();
  return e;
}

// ------------------------------------------------------------------------------------------
// raw (non-differentiable) implementations
// ------------------------------------------------------------------------------------------
Tensor align_fwd_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x) {
  check_x(x, "align");
  NvtxRange nvtx("molann_b200::align");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  TORCH_CHECK(h.plan.n_align > 0, "molann_b200::align: empty alignment selection");
  Tensor out = at::empty_like(x);
  check_status(molann_b200_align_forward(&h.plan, x.data_ptr<float>(), x.size(0), out.data_ptr<float>(), cur_stream()),
               "align_forward");
  return out;
}

Tensor align_bwd_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& gout_in) {
  check_cotangent(gout_in, x, "align backward");
  NvtxRange nvtx("molann_b200::align backward");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  Tensor gout = gout_in.contiguous();
  Tensor gx = at::empty_like(x);
  check_status(molann_b200_align_backward(&h.plan, x.data_ptr<float>(), gout.data_ptr<float>(), x.size(0),
                                          gx.data_ptr<float>(), cur_stream()),
               "align_backward");
  return gx;
}

Tensor preprocess_fwd_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                           int64_t d_feat, bool use_angle_value) {
  check_x(x, "preprocess");
  NvtxRange nvtx("molann_b200::preprocess");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  Tensor feat = at::empty({x.size(0), d_feat}, x.options());
  check_status(molann_b200_preprocess_forward(&h.plan, x.data_ptr<float>(), x.size(0), feat.data_ptr<float>(),
                                              cur_stream()),
               "preprocess_forward");
  return feat;
}

Tensor preprocess_bwd_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                           int64_t d_feat, bool use_angle_value, const Tensor& gfeat_in) {
  check_cotangent(gfeat_in, x, "preprocess backward");
  NvtxRange nvtx("molann_b200::preprocess backward");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  Tensor gfeat = gfeat_in.contiguous();
  Tensor gx = at::empty_like(x);
  check_status(molann_b200_preprocess_backward(&h.plan, x.data_ptr<float>(), gfeat.data_ptr<float>(), x.size(0),
                                               gx.data_ptr<float>(), cur_stream()),
               "preprocess_backward");
  return gx;
}

Tensor molann_fwd_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                       int64_t d_feat, bool use_angle_value, at::TensorList params, int64_t act) {
  check_x(x, "molann");
  NvtxRange nvtx("molann_b200::molann");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  fill_mlp(h, x, params, act);
  const int64_t L = x.size(0);
  Tensor y = at::empty({L, h.plan.dims[h.plan.n_layers]}, x.options());
  Tensor ws;
  void* wsp = nullptr;
  size_t ws_bytes = 0;
  if (L > 0 && molann_b200_wide_eligible(&h.plan)) {
    // big system, wide first layer: ONE persistent kernel on a prepared plan (csrc/fused_wide.cuh)
    auto prep = prepared_for(h, x, entries, align_idx, ref_x, params, act, use_angle_value);
    if (!prep->unsupported) {
      ws_bytes = molann_b200_prepared_workspace_bytes(prep->handle, L);
      ws = at::empty({static_cast<int64_t>(ws_bytes)}, x.options().dtype(at::kByte));
      const int st = molann_b200_forward_prepared(prep->handle, &h.plan, x.data_ptr<float>(), L, y.data_ptr<float>(),
                                                  ws.data_ptr(), ws_bytes, cur_stream());
      if (st != MOLANN_ERR_UNSUPPORTED) {
        check_status(st, "forward_prepared");
        return y;
      }
      prep->unsupported = true;          // e.g. frames too big for the shared-memory ring: the layered kernels serve it
    }
  }
  if (molann_b200_path_for(&h.plan, 0) != 1) {
    ws_bytes = molann_b200_workspace_bytes(&h.plan, L, 0);
    ws = at::empty({static_cast<int64_t>(ws_bytes)}, x.options().dtype(at::kByte));
    wsp = ws.data_ptr();
  }
  check_status(molann_b200_forward(&h.plan, x.data_ptr<float>(), L, y.data_ptr<float>(), wsp, ws_bytes, cur_stream()),
               "forward");
  return y;
}

// returns {gx, gW1, gb1, gW2, gb2, ...}; parameter gradients are undefined tensors unless requested
std::vector<Tensor> molann_bwd_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x,
                                    const Tensor& entries, int64_t d_feat, bool use_angle_value,
                                    at::TensorList params, int64_t act, const Tensor& gy_in, bool want_params,
                                    bool want_x = true) {
  check_cotangent(gy_in, x, "molann backward");
  NvtxRange nvtx("molann_b200::molann backward");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  fill_mlp(h, x, params, act);
  const int64_t L = x.size(0);
  const int nl = h.plan.n_layers;
  Tensor gy = gy_in.contiguous();
  want_x = want_x || !want_params;
  Tensor gx = want_x ? at::empty_like(x) : Tensor();
  std::vector<Tensor> out;
  out.push_back(gx);
  float* gW[MOLANN_MAX_LAYERS] = {nullptr};
  float* gb[MOLANN_MAX_LAYERS] = {nullptr};
  if (want_params) {
    for (int k = 0; k < nl; ++k) {
      out.push_back(at::zeros_like(params[2 * k], at::MemoryFormat::Contiguous));
      out.push_back(at::zeros_like(params[2 * k + 1], at::MemoryFormat::Contiguous));
      gW[k] = out[1 + 2 * k].data_ptr<float>();
      gb[k] = out[2 + 2 * k].data_ptr<float>();
    }
  }
  Tensor ws;
  void* wsp = nullptr;
  size_t ws_bytes = 0;
  if (want_params || molann_b200_path_for(&h.plan, 1) != 1) {
    ws_bytes = molann_b200_workspace_bytes(&h.plan, L, 1);
    ws = at::empty({static_cast<int64_t>(ws_bytes)}, x.options().dtype(at::kByte));
    wsp = ws.data_ptr();
  }
  check_status(molann_b200_backward(&h.plan, x.data_ptr<float>(), gy.data_ptr<float>(), L,
                                    want_x ? gx.data_ptr<float>() : nullptr,
                                    want_params ? gW : nullptr, want_params ? gb : nullptr, wsp, ws_bytes,
                                    cur_stream()),
               "backward");
  return out;
}

// y = model(x) and gx = d<gy, y>/dx in ONE pass (biasing-force entry point, no autograd graph)
std::tuple<Tensor, Tensor> value_and_grad_impl(const Tensor& x, const Tensor& gy_in, const Tensor& align_idx,
                                               const Tensor& ref_x, const Tensor& entries, int64_t d_feat,
                                               bool use_angle_value, at::TensorList params, int64_t act) {
  check_x(x, "value_and_grad");
  NvtxRange nvtx("molann_b200::value_and_grad");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  fill_mlp(h, x, params, act);
  const int64_t L = x.size(0);
  const int64_t kout = h.plan.dims[h.plan.n_layers];
  check_cotangent(gy_in, x, "value_and_grad");
  TORCH_CHECK(gy_in.dim() == 2 && gy_in.size(0) == L && gy_in.size(1) == kout,
              "molann_b200::value_and_grad: cotangent must have shape [L, ", kout, "]");
  Tensor gy = gy_in.contiguous();
  Tensor y = at::empty({L, kout}, x.options());
  Tensor gx = at::empty_like(x);
  Tensor ws;
  void* wsp = nullptr;
  size_t ws_bytes = 0;
  if (L > 0 && molann_b200_wide_eligible(&h.plan)) {           // big system: packed operands from the prepared plan
    auto prep = prepared_for(h, x, entries, align_idx, ref_x, params, act, use_angle_value);
    ws_bytes = molann_b200_prepared_workspace_bytes(prep->handle, L);
    ws = at::empty({static_cast<int64_t>(ws_bytes)}, x.options().dtype(at::kByte));
    check_status(molann_b200_value_and_grad_prepared(prep->handle, &h.plan, x.data_ptr<float>(), gy.data_ptr<float>(),
                                                     L, y.data_ptr<float>(), gx.data_ptr<float>(), ws.data_ptr(),
                                                     ws_bytes, cur_stream()),
                 "value_and_grad_prepared");
    return std::make_tuple(y, gx);
  }
  if (molann_b200_kernel_family(&h.plan, 1) != 2) {
    ws_bytes = molann_b200_workspace_bytes(&h.plan, L, 1);
    ws = at::empty({static_cast<int64_t>(ws_bytes)}, x.options().dtype(at::kByte));
    wsp = ws.data_ptr();
  }
  check_status(molann_b200_value_and_grad(&h.plan, x.data_ptr<float>(), gy.data_ptr<float>(), L, y.data_ptr<float>(),
                                          gx.data_ptr<float>(), wsp, ws_bytes, cur_stream()),
               "value_and_grad");
  return std::make_tuple(y, gx);
}

// y = model(x) and jac[k, L, n_inp, 3] = d y[:, o] / dx for every output o (no autograd graph)
std::tuple<Tensor, Tensor> value_and_jacobian_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x,
                                                   const Tensor& entries, int64_t d_feat, bool use_angle_value,
                                                   at::TensorList params, int64_t act) {
  check_x(x, "value_and_jacobian");
  NvtxRange nvtx("molann_b200::value_and_jacobian");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  fill_mlp(h, x, params, act);
  const int64_t L = x.size(0);
  const int64_t kout = h.plan.dims[h.plan.n_layers];
  Tensor y = at::empty({L, kout}, x.options());
  Tensor jac = at::empty({kout, L, x.size(1), 3}, x.options());
  const size_t ws_bytes = molann_b200_jacobian_workspace_bytes(&h.plan, L);
  Tensor ws = at::empty({static_cast<int64_t>(ws_bytes > 0 ? ws_bytes : 1)}, x.options().dtype(at::kByte));
  check_status(molann_b200_value_and_jacobian(&h.plan, x.data_ptr<float>(), L, y.data_ptr<float>(),
                                              jac.data_ptr<float>(), ws.data_ptr(), ws_bytes, cur_stream()),
               "value_and_jacobian");
  return std::make_tuple(y, jac);
}

// ------------------------------------------------------------------------------------------
// autograd
// ------------------------------------------------------------------------------------------
struct AlignFn : public torch::autograd::Function<AlignFn> {
  static Tensor forward(AutogradContext* ctx, const Tensor& x, const Tensor& align_idx, const Tensor& ref_x) {
    at::AutoDispatchBelowADInplaceOrView g;
    ctx->save_for_backward({x, align_idx, ref_x});
    return align_fwd_impl(x, align_idx, ref_x);
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    refuse_double_backward("align");
    auto saved = ctx->get_saved_variables();
    Tensor gx = align_bwd_impl(saved[0], saved[1], saved[2], grads[0]);
    return {gx, Tensor(), Tensor()};
  }
};

struct PreprocessFn : public torch::autograd::Function<PreprocessFn> {
  static Tensor forward(AutogradContext* ctx, const Tensor& x, const Tensor& align_idx, const Tensor& ref_x,
                        const Tensor& entries, int64_t d_feat, bool use_angle_value) {
    at::AutoDispatchBelowADInplaceOrView g;
    ctx->save_for_backward({x, align_idx, ref_x, entries});
    ctx->saved_data["d_feat"] = d_feat;
    ctx->saved_data["use_angle_value"] = use_angle_value;
    return preprocess_fwd_impl(x, align_idx, ref_x, entries, d_feat, use_angle_value);
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    refuse_double_backward("preprocess");
    auto saved = ctx->get_saved_variables();
    Tensor gx = preprocess_bwd_impl(saved[0], saved[1], saved[2], saved[3], ctx->saved_data["d_feat"].toInt(),
                                    ctx->saved_data["use_angle_value"].toBool(), grads[0]);
    return {gx, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

struct MolannFn : public torch::autograd::Function<MolannFn> {
  static Tensor forward(AutogradContext* ctx, const Tensor& x, const Tensor& align_idx, const Tensor& ref_x,
                        const Tensor& entries, int64_t d_feat, bool use_angle_value, at::TensorList params,
                        int64_t act) {
    at::AutoDispatchBelowADInplaceOrView g;
    std::vector<Tensor> to_save = {x, align_idx, ref_x, entries};
    bool want_params = false;
    for (const Tensor& p : params) {
      to_save.push_back(p);
      want_params = want_params || p.requires_grad();
    }
    ctx->save_for_backward(to_save);
    ctx->saved_data["d_feat"] = d_feat;
    ctx->saved_data["use_angle_value"] = use_angle_value;
    ctx->saved_data["act"] = act;
    ctx->saved_data["want_params"] = want_params;
    ctx->saved_data["n_params"] = static_cast<int64_t>(params.size());
    return molann_fwd_impl(x, align_idx, ref_x, entries, d_feat, use_angle_value, params, act);
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    refuse_double_backward("molann");
    auto saved = ctx->get_saved_variables();
    const int64_t np = ctx->saved_data["n_params"].toInt();
    std::vector<Tensor> params(saved.begin() + 4, saved.begin() + 4 + np);
    // parameter gradients only when the engine actually asks for them (torch.autograd.grad(y, x) for the
    // biasing-force path does not, even if the Linear parameters have requires_grad=True)
    bool want_params = false;
    if (ctx->saved_data["want_params"].toBool())
      for (int64_t i = 0; i < np; ++i) want_params = want_params || ctx->needs_input_grad(4 + i);
    auto g = molann_bwd_impl(saved[0], saved[1], saved[2], saved[3], ctx->saved_data["d_feat"].toInt(),
                             ctx->saved_data["use_angle_value"].toBool(), params, ctx->saved_data["act"].toInt(),
                             grads[0], want_params, ctx->needs_input_grad(0));
    variable_list out;
    out.push_back(g[0]);                                         // x (undefined when x needs no gradient)
    for (int i = 0; i < 5; ++i) out.push_back(Tensor());         // align_idx, ref_x, entries, d_feat, use_angle
    for (int64_t i = 0; i < np; ++i) out.push_back(want_params ? g[1 + i] : Tensor());
    out.push_back(Tensor());                                     // act
    return out;
  }
};

Tensor align_autograd(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x) {
  check_x(x, "align");
  return AlignFn::apply(x, align_idx, ref_x);
}
Tensor preprocess_autograd(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                           int64_t d_feat, bool use_angle_value) {
  check_x(x, "preprocess");
  return PreprocessFn::apply(x, align_idx, ref_x, entries, d_feat, use_angle_value);
}
Tensor molann_autograd(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                       int64_t d_feat, bool use_angle_value, at::TensorList params, int64_t act) {
  check_x(x, "molann");
  return MolannFn::apply(x, align_idx, ref_x, entries, d_feat, use_angle_value, params, act);
}

// value_and_grad is an explicit "no graph" entry point: its outputs never carry autograd history, whatever the
// requires_grad flags of x and the MLP parameters say (without this kernel PyTorch's autograd-not-implemented
// fallback would hand back tensors with requires_grad=True and a grad_fn that only warns).
std::tuple<Tensor, Tensor> value_and_grad_autograd(const Tensor& x, const Tensor& gy, const Tensor& align_idx,
                                                   const Tensor& ref_x, const Tensor& entries, int64_t d_feat,
                                                   bool use_angle_value, at::TensorList params, int64_t act) {
  // (detach BEFORE the guard: below ADInplaceOrView a detach makes a fresh version counter, and the prepared-plan cache
  // keys on the parameters' versions)
  std::vector<Tensor> plain;
  for (const Tensor& p : params) plain.push_back(p.detach());
  const Tensor xd = x.detach(), gyd = gy.detach();
  at::AutoDispatchBelowADInplaceOrView g;
  return value_and_grad_impl(xd, gyd, align_idx, ref_x, entries, d_feat, use_angle_value, plain, act);
}

std::tuple<Tensor, Tensor> value_and_jacobian_autograd(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x,
                                                       const Tensor& entries, int64_t d_feat, bool use_angle_value,
                                                       at::TensorList params, int64_t act) {
  std::vector<Tensor> plain;
  for (const Tensor& p : params) plain.push_back(p.detach());
  const Tensor xd = x.detach();
  at::AutoDispatchBelowADInplaceOrView g;
  return value_and_jacobian_impl(xd, align_idx, ref_x, entries, d_feat, use_angle_value, plain, act);
}

// int16 wire frames -> fp32 coordinates (trajectory ingestion, include/molann_b200.h)
Tensor decode_frames_impl(const Tensor& q, double ox, double oy, double oz, double resolution) {
  TORCH_CHECK(q.is_cuda(), "molann_b200::decode_frames: the wire frames must be a CUDA tensor, got ", q.device());
  TORCH_CHECK(q.scalar_type() == at::kShort, "molann_b200::decode_frames: wire frames must be int16, got ",
              q.scalar_type());
  TORCH_CHECK(q.dim() == 3 && q.size(2) == 3 && q.is_contiguous(),
              "molann_b200::decode_frames: wire frames must be a contiguous [L, n_inp, 3] tensor");
  NvtxRange nvtx("molann_b200::decode_frames");
  c10::cuda::CUDAGuard guard(q.device());
  Tensor x = at::empty(q.sizes(), q.options().dtype(at::kFloat));
  const float origin[3] = {static_cast<float>(ox), static_cast<float>(oy), static_cast<float>(oz)};
  check_status(molann_b200_decode_frames_i16(q.data_ptr<int16_t>(), q.numel(), origin, static_cast<float>(resolution),
                                             x.data_ptr<float>(), cur_stream()),
               "decode_frames_i16");
  return x;
}

// ---- autoencoder training step (include/molann_b200.h: molann_b200_train_*) ----
struct DecoderHolder {
  MolannDecoder dec;
  std::vector<Tensor> keep;
};

void fill_decoder(DecoderHolder& d, const PlanHolder& h, const Tensor& x, at::TensorList params, int64_t act) {
  std::memset(&d.dec, 0, sizeof(MolannDecoder));
  TORCH_CHECK(params.size() >= 2 && params.size() % 2 == 0, "molann_b200: decoder params must be [W1, b1, W2, b2, ...]");
  const int nl = static_cast<int>(params.size() / 2);
  TORCH_CHECK(nl <= MOLANN_MAX_LAYERS, "molann_b200: at most ", MOLANN_MAX_LAYERS, " decoder layers are supported");
  d.dec.n_layers = nl;
  d.dec.act_id = static_cast<int32_t>(act);
  d.dec.dims[0] = h.plan.dims[h.plan.n_layers];
  for (int k = 0; k < nl; ++k) {
    const Tensor& W = params[2 * k];
    const Tensor& b = params[2 * k + 1];
    TORCH_CHECK(W.dim() == 2 && b.dim() == 1 && W.size(0) == b.size(0), "molann_b200: bad decoder Linear shapes at layer ",
                k + 1);
    TORCH_CHECK(W.size(1) == d.dec.dims[k], "molann_b200: decoder layer ", k + 1, " expects ", W.size(1),
                " inputs but receives ", d.dec.dims[k]);
    TORCH_CHECK(W.device() == x.device() && b.device() == x.device(),
                "molann_b200: decoder parameters must live on the input's device -- call module.to(device)");
    TORCH_CHECK(W.scalar_type() == at::kFloat && b.scalar_type() == at::kFloat,
                "molann_b200: decoder parameters must be float32");
    d.keep.push_back(W.contiguous());
    d.keep.push_back(b.contiguous());
    d.dec.W[k] = d.keep[2 * k].data_ptr<float>();
    d.dec.b[k] = d.keep[2 * k + 1].data_ptr<float>();
    d.dec.dims[k + 1] = static_cast<int32_t>(W.size(0));
  }
  TORCH_CHECK(d.dec.dims[nl] == h.plan.d_feat, "molann_b200: the decoder must reconstruct the ", h.plan.d_feat,
              " features, its last layer has ", d.dec.dims[nl], " outputs");
}

bool train_eligible_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                         int64_t d_feat, bool use_angle_value, at::TensorList enc_params, int64_t enc_act,
                         at::TensorList dec_params, int64_t dec_act) {
  check_x(x, "train_eligible");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  fill_mlp(h, x, enc_params, enc_act);
  DecoderHolder d;
  fill_decoder(d, h, x, dec_params, dec_act);
  return molann_b200_train_eligible(&h.plan, &d.dec) != 0;
}

// flat[P + 1]: d loss / d (encoder, decoder parameters) in parameter order, then the loss (see the header)
Tensor train_loss_and_grads_impl(const Tensor& x, const Tensor& align_idx, const Tensor& ref_x, const Tensor& entries,
                                 int64_t d_feat, bool use_angle_value, at::TensorList enc_params, int64_t enc_act,
                                 at::TensorList dec_params, int64_t dec_act, double loss_scale) {
  check_x(x, "train_loss_and_grads");
  NvtxRange nvtx("molann_b200::train_loss_and_grads");
  c10::cuda::CUDAGuard guard(x.device());
  PlanHolder h;
  fill_geometry(h, x, align_idx, ref_x);
  fill_features(h, x, entries, d_feat, use_angle_value);
  fill_mlp(h, x, enc_params, enc_act);
  DecoderHolder d;
  fill_decoder(d, h, x, dec_params, dec_act);
  const size_t P = molann_b200_train_param_count(&h.plan, &d.dec);
  TORCH_CHECK(P > 0, "molann_b200::train_loss_and_grads: inconsistent encoder / decoder");
  Tensor flat = at::empty({static_cast<int64_t>(P) + 1}, x.options());
  const size_t ws_bytes = molann_b200_train_workspace_bytes(&h.plan, &d.dec);
  Tensor ws = at::empty({static_cast<int64_t>(ws_bytes)}, x.options().dtype(at::kByte));
  check_status(molann_b200_train_loss_and_grads(&h.plan, &d.dec, x.data_ptr<float>(), x.size(0),
                                                static_cast<float>(loss_scale), flat.data_ptr<float>(), ws.data_ptr(),
                                                ws_bytes, cur_stream()),
               "train_loss_and_grads");
  return flat;
}

// params[i] -= lr * flat[offset_i : offset_i + numel_i]   (in place, one launch)
void sgd_apply_impl(at::TensorList params, const Tensor& flat, double lr) {
  TORCH_CHECK(flat.is_cuda() && flat.scalar_type() == at::kFloat && flat.is_contiguous(),
              "molann_b200::sgd_apply_: the flat gradient must be a contiguous float32 CUDA tensor");
  TORCH_CHECK(params.size() <= 4 * MOLANN_MAX_LAYERS, "molann_b200::sgd_apply_: too many parameter tensors");
  c10::cuda::CUDAGuard guard(flat.device());
  std::vector<float*> ptrs;
  std::vector<int64_t> numel;
  int64_t total = 0;
  for (const Tensor& p : params) {
    TORCH_CHECK(p.device() == flat.device() && p.scalar_type() == at::kFloat && p.is_contiguous(),
                "molann_b200::sgd_apply_: parameters must be contiguous float32 tensors on the gradient's device");
    ptrs.push_back(p.data_ptr<float>());
    numel.push_back(p.numel());
    total += p.numel();
  }
  TORCH_CHECK(total <= flat.numel(), "molann_b200::sgd_apply_: the flat gradient holds ", flat.numel(),
              " values, the parameters need ", total);
  check_status(molann_b200_sgd_apply(ptrs.data(), numel.data(), static_cast<int32_t>(ptrs.size()),
                                     flat.data_ptr<float>(), static_cast<float>(lr), cur_stream()),
               "sgd_apply");
}

// one-shot allreduce of the flat vector over peer memory + SGD (include/molann_b200.h: molann_b200_allreduce_sgd).
// `peer_ptrs[r]` = rank r's symmetric buffer as seen from this device (torch.distributed._symmetric_memory buffer_ptrs).
Tensor allreduce_sgd_impl(at::TensorList params, const Tensor& flat_local, at::IntArrayRef peer_ptrs, int64_t rank,
                          const Tensor& state, double lr) {
  TORCH_CHECK(flat_local.is_cuda() && flat_local.scalar_type() == at::kFloat && flat_local.is_contiguous(),
              "molann_b200::allreduce_sgd_: the flat vector must be a contiguous float32 CUDA tensor");
  TORCH_CHECK(state.is_cuda() && state.device() == flat_local.device() && state.scalar_type() == at::kInt &&
                  state.numel() >= 3 && state.is_contiguous(),
              "molann_b200::allreduce_sgd_: state must be an int32[3] tensor on the vector's device");
  TORCH_CHECK(!peer_ptrs.empty() && peer_ptrs.size() <= 8, "molann_b200::allreduce_sgd_: 1 .. 8 peers");
  c10::cuda::CUDAGuard guard(flat_local.device());
  NvtxRange nvtx("molann_b200::allreduce_sgd_");
  std::vector<float*> ptrs;
  std::vector<int64_t> numel;
  for (const Tensor& p : params) {
    TORCH_CHECK(p.device() == flat_local.device() && p.scalar_type() == at::kFloat && p.is_contiguous(),
                "molann_b200::allreduce_sgd_: parameters must be contiguous float32 tensors on the vector's device");
    ptrs.push_back(p.data_ptr<float>());
    numel.push_back(p.numel());
  }
  std::vector<void*> bufs;
  for (int64_t a : peer_ptrs) bufs.push_back(reinterpret_cast<void*>(static_cast<uintptr_t>(a)));
  Tensor out = at::empty_like(flat_local);
  check_status(molann_b200_allreduce_sgd(flat_local.data_ptr<float>(), out.data_ptr<float>(), flat_local.numel(),
                                         bufs.data(), static_cast<int32_t>(rank), static_cast<int32_t>(bufs.size()),
                                         reinterpret_cast<uint32_t*>(state.data_ptr<int32_t>()),
                                         ptrs.empty() ? nullptr : ptrs.data(), numel.empty() ? nullptr : numel.data(),
                                         static_cast<int32_t>(ptrs.size()), static_cast<float>(lr), cur_stream()),
               "allreduce_sgd");
  return out;
}

int64_t allreduce_buffer_bytes(int64_t n, int64_t world) {
  return static_cast<int64_t>(molann_b200_allreduce_buffer_bytes(n, static_cast<int32_t>(world)));
}

int64_t launch_count() { return molann_b200_launch_count(); }

}  // namespace

TORCH_LIBRARY(molann_b200, m) {
  m.def("align(Tensor x, Tensor align_idx, Tensor ref_x) -> Tensor");
  m.def("preprocess(Tensor x, Tensor align_idx, Tensor ref_x, Tensor entries, int d_feat, bool use_angle_value) -> Tensor");
  m.def("molann(Tensor x, Tensor align_idx, Tensor ref_x, Tensor entries, int d_feat, bool use_angle_value, "
        "Tensor[] params, int act) -> Tensor");
  m.def("value_and_grad(Tensor x, Tensor gy, Tensor align_idx, Tensor ref_x, Tensor entries, int d_feat, "
        "bool use_angle_value, Tensor[] params, int act) -> (Tensor, Tensor)");
  m.def("value_and_jacobian(Tensor x, Tensor align_idx, Tensor ref_x, Tensor entries, int d_feat, "
        "bool use_angle_value, Tensor[] params, int act) -> (Tensor, Tensor)");
  m.def("decode_frames(Tensor q, float ox, float oy, float oz, float resolution) -> Tensor");
  m.def("train_eligible(Tensor x, Tensor align_idx, Tensor ref_x, Tensor entries, int d_feat, bool use_angle_value, "
        "Tensor[] enc_params, int enc_act, Tensor[] dec_params, int dec_act) -> bool");
  m.def("train_loss_and_grads(Tensor x, Tensor align_idx, Tensor ref_x, Tensor entries, int d_feat, "
        "bool use_angle_value, Tensor[] enc_params, int enc_act, Tensor[] dec_params, int dec_act, float loss_scale) "
        "-> Tensor");
  m.def("sgd_apply_(Tensor(a!)[] params, Tensor flat, float lr) -> ()");
  m.def("allreduce_sgd_(Tensor(a!)[] params, Tensor flat_local, int[] peer_ptrs, int rank, Tensor(b!) state, float lr) "
        "-> Tensor");
  m.def("allreduce_buffer_bytes(int n, int world) -> int", &allreduce_buffer_bytes);
  m.def("launch_count() -> int", &launch_count);
}

TORCH_LIBRARY_IMPL(molann_b200, CUDA, m) {
  m.impl("align", &align_fwd_impl);
  m.impl("preprocess", &preprocess_fwd_impl);
  m.impl("molann", &molann_fwd_impl);
  m.impl("value_and_grad", &value_and_grad_impl);
  m.impl("value_and_jacobian", &value_and_jacobian_impl);
  m.impl("decode_frames", &decode_frames_impl);
  m.impl("train_eligible", &train_eligible_impl);
  m.impl("train_loss_and_grads", &train_loss_and_grads_impl);
  m.impl("sgd_apply_", &sgd_apply_impl);
  m.impl("allreduce_sgd_", &allreduce_sgd_impl);
}

TORCH_LIBRARY_IMPL(molann_b200, Autograd, m) {
  m.impl("align", &align_autograd);
  m.impl("preprocess", &preprocess_autograd);
  m.impl("molann", &molann_autograd);
  m.impl("value_and_grad", &value_and_grad_autograd);
  m.impl("value_and_jacobian", &value_and_jacobian_autograd);
}
