// consumer.cpp -- a libtorch-only consumer of an exported molann_b200 model, the way an MD-engine plugin uses the
// reference's TorchScript files (README.rst:51, molann/ann.py:109-111): no Python, no molann_b200 package -- it
// dlopen()s the custom-op shim, torch::jit::load()s model.pt, and evaluates the collective variables, their
// coordinate gradient through C++ autograd, and the exported value_and_grad / value_and_jacobian entry points.
//
//   consumer <libmolann_b200_torch.so> <model.pt> <io.pt>
// io.pt is a scripted container module with buffers x, cot and the values Python computed (y, gx, jac).
#include <dlfcn.h>
#include <torch/script.h>
#include <torch/torch.h>

#include <cstdio>
#include <string>

static double max_abs_diff(const torch::Tensor& a, const torch::Tensor& b) {
  return (a.to(torch::kCPU).to(torch::kDouble) - b.to(torch::kCPU).to(torch::kDouble)).abs().max().item<double>();
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: consumer <shim.so> <model.pt> <io.pt>\n");
    return 2;
  }
  if (!dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL)) {               // registers torch.ops.molann_b200.*
    std::fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 3;
  }
  if (!torch::cuda::is_available()) {
    std::fprintf(stderr, "no CUDA device: molann_b200 has no CPU path\n");
    return 4;
  }
  torch::jit::script::Module model = torch::jit::load(argv[2], torch::kCUDA);
  torch::jit::script::Module io = torch::jit::load(argv[3], torch::kCPU);
  const torch::Tensor x = io.attr("x").toTensor().to(torch::kCUDA);
  const torch::Tensor cot = io.attr("cot").toTensor().to(torch::kCUDA);
  const torch::Tensor y_ref = io.attr("y").toTensor(), gx_ref = io.attr("gx").toTensor(), jac_ref = io.attr("jac").toTensor();

  // 1. forward + C++ autograd (the custom op registers its backward in C++, torch_shim.cpp)
  torch::Tensor xg = x.clone().set_requires_grad(true);
  torch::Tensor y = model.forward({xg}).toTensor();
  auto grads = torch::autograd::grad({y}, {xg}, {cot});
  const double ey = max_abs_diff(y.detach(), y_ref), eg = max_abs_diff(grads[0], gx_ref);

  // 2. the fused biasing-force entry points exported with the model
  auto vg = model.get_method("value_and_grad")({x, cot}).toTuple();
  const double ey2 = max_abs_diff(vg->elements()[0].toTensor(), y_ref);
  const double eg2 = max_abs_diff(vg->elements()[1].toTensor(), gx_ref);
  auto vj = model.get_method("value_and_jacobian")({x}).toTuple();
  const double ey3 = max_abs_diff(vj->elements()[0].toTensor(), y_ref);
  const double ej = max_abs_diff(vj->elements()[1].toTensor(), jac_ref);
  torch::cuda::synchronize();

  std::printf("consumer: L=%lld  |dy| %.3e  |dgx| %.3e  value_and_grad |dy| %.3e |dgx| %.3e  value_and_jacobian |dy| %.3e |dJ| %.3e\n",
              (long long)x.size(0), ey, eg, ey2, eg2, ey3, ej);
  const double scale = gx_ref.abs().max().item<double>();
  const bool ok = ey == 0.0 && eg == 0.0 && ey2 <= 2e-6 && eg2 <= 2e-5 * scale && ey3 <= 2e-6 && ej <= 2e-5 * scale;
  std::printf(ok ? "CONSUMER_OK\n" : "CONSUMER_MISMATCH\n");
  return ok ? 0 : 1;
}
