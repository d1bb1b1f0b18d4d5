"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  libmolann_b200.so        hand-written sm_100a kernels + the C ABI of include/molann_b200.h  (nvcc)
  libmolann_b200_torch.so  TORCH_LIBRARY shim + C++ autograd over that C ABI                  (g++)

Run ``python -m molann_b200.build`` or call :func:`build_all`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_KERNELS = os.path.join(HERE, "libmolann_b200.so")
LIB_TORCH = os.path.join(HERE, "libmolann_b200_torch.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA toolkit is required to build molann_b200")


def _newer(target, sources):
    if not os.path.isfile(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources():
    inc = os.path.join(os.path.dirname(HERE), "include", "molann_b200.h")
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    return cu + [inc]


def build_kernels(force=False, verbose=False):
    srcs = _sources()
    if not force and _newer(LIB_KERNELS, srcs):
        return LIB_KERNELS
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [_nvcc()] + flags + ["-shared", "-o", LIB_KERNELS, os.path.join(CSRC, "molann_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_KERNELS


def build_torch_shim(force=False, verbose=False):
    import torch
    from torch.utils import cpp_extension
    src = os.path.join(CSRC, "torch_shim.cpp")
    if not force and _newer(LIB_TORCH, [src, LIB_KERNELS] + _sources()):
        return LIB_TORCH
    inc = []
    for p in cpp_extension.include_paths():
        inc += ["-isystem", p]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc += ["-isystem", os.path.join(cuda_home, "include")]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
           src, "-o", LIB_TORCH] + inc + [
        "-L" + torch_lib, "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda",
        "-L" + HERE, "-lmolann_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib, "-Wl,--no-as-needed"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_TORCH


CONSUMER_SRC = os.path.join(os.path.dirname(HERE), "tests", "host", "consumer.cpp")
CONSUMER_BIN = os.path.join(os.path.dirname(HERE), "tests", "host", "consumer")


def build_consumer(force=False, verbose=False):
    """tests/host/consumer: the libtorch-only C++ program that loads an exported model the way an MD plugin does."""
    import torch
    from torch.utils import cpp_extension
    if not force and _newer(CONSUMER_BIN, [CONSUMER_SRC]):
        return CONSUMER_BIN
    inc = []
    for p in cpp_extension.include_paths():
        inc += ["-isystem", p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O1", "-std=c++17", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi, CONSUMER_SRC, "-o", CONSUMER_BIN] + inc + [
        "-L" + torch_lib, "-Wl,--no-as-needed", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda",
        "-ldl", "-Wl,-rpath," + torch_lib]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return CONSUMER_BIN


def build_all(force=False, verbose=False):
    build_kernels(force=force, verbose=verbose)
    build_torch_shim(force=force, verbose=verbose)
    return LIB_KERNELS, LIB_TORCH


if __name__ == "__main__":
    libs = build_all(force="--force" in sys.argv, verbose=True)
    print("built:", *libs)
