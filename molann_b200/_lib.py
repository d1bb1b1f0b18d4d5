"""Loader for the native libraries.  Fails loudly: there is no Python / CPU fallback.

``cabi()``          ctypes handle on libmolann_b200.so (the C ABI of include/molann_b200.h)
``load_torch_ops()`` registers ``torch.ops.molann_b200.*`` (libmolann_b200_torch.so)
"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_KERNELS = os.path.join(HERE, "libmolann_b200.so")
LIB_TORCH = os.path.join(HERE, "libmolann_b200_torch.so")
MAX_LAYERS = 8
ENTRY_INTS = 6

_HELP = ("molann_b200: native library '%s' is missing or failed to load (%s). Build it in-tree with "
         "`python -m molann_b200.build` (needs nvcc; target sm_100a). There is no CPU fallback.")


class MolannPlan(ctypes.Structure):
    """Mirror of ``struct MolannPlan`` (include/molann_b200.h)."""
    _fields_ = [
        ("n_inp", ctypes.c_int32), ("n_align", ctypes.c_int32),
        ("align_idx", ctypes.c_void_p), ("ref_x", ctypes.c_void_p),
        ("n_entries", ctypes.c_int32), ("entries", ctypes.c_void_p),
        ("d_feat", ctypes.c_int32), ("use_angle_value", ctypes.c_int32),
        ("n_layers", ctypes.c_int32), ("act_id", ctypes.c_int32),
        ("dims", ctypes.c_int32 * (MAX_LAYERS + 1)),
        ("W", ctypes.c_void_p * MAX_LAYERS), ("b", ctypes.c_void_p * MAX_LAYERS),
    ]


class MolannDecoder(ctypes.Structure):
    """Mirror of ``struct MolannDecoder`` (include/molann_b200.h)."""
    _fields_ = [
        ("n_layers", ctypes.c_int32), ("act_id", ctypes.c_int32),
        ("dims", ctypes.c_int32 * (MAX_LAYERS + 1)),
        ("W", ctypes.c_void_p * MAX_LAYERS), ("b", ctypes.c_void_p * MAX_LAYERS),
    ]


_cabi = None
_ops_loaded = False


def cabi():
    global _cabi
    if _cabi is not None:
        return _cabi
    if not os.path.isfile(LIB_KERNELS):
        raise RuntimeError(_HELP % (LIB_KERNELS, "file not found"))
    try:
        lib = ctypes.CDLL(LIB_KERNELS)
    except OSError as exc:
        raise RuntimeError(_HELP % (LIB_KERNELS, exc)) from exc
    P = ctypes.POINTER(MolannPlan)
    vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
    lib.molann_b200_version.restype = ctypes.c_int
    lib.molann_b200_strerror.restype = ctypes.c_char_p
    lib.molann_b200_strerror.argtypes = [ctypes.c_int]
    lib.molann_b200_last_cuda_error.restype = ctypes.c_int
    lib.molann_b200_cuda_error_string.restype = ctypes.c_char_p
    lib.molann_b200_cuda_error_string.argtypes = [ctypes.c_int]
    lib.molann_b200_launch_count.restype = i64
    lib.molann_b200_plan_validate.argtypes = [P]
    lib.molann_b200_path_for.argtypes = [P, ctypes.c_int]
    lib.molann_b200_kernel_family.argtypes = [P, ctypes.c_int]
    lib.molann_b200_workspace_bytes.restype = sz
    lib.molann_b200_workspace_bytes.argtypes = [P, i64, ctypes.c_int]
    lib.molann_b200_forward.argtypes = [P, vp, i64, vp, vp, sz, vp]
    lib.molann_b200_backward.argtypes = [P, vp, vp, i64, vp, vp, vp, vp, sz, vp]
    lib.molann_b200_value_and_grad.argtypes = [P, vp, vp, i64, vp, vp, vp, sz, vp]
    lib.molann_b200_preprocess_forward.argtypes = [P, vp, i64, vp, vp]
    lib.molann_b200_preprocess_backward.argtypes = [P, vp, vp, i64, vp, vp]
    lib.molann_b200_align_forward.argtypes = [P, vp, i64, vp, vp]
    lib.molann_b200_align_backward.argtypes = [P, vp, vp, i64, vp, vp]
    lib.molann_b200_decode_frames_i16.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_float), ctypes.c_float, vp, vp]
    lib.molann_b200_jacobian_workspace_bytes.restype = sz
    lib.molann_b200_jacobian_workspace_bytes.argtypes = [P, i64]
    lib.molann_b200_value_and_jacobian.argtypes = [P, vp, i64, vp, vp, vp, sz, vp]
    lib.molann_b200_wide_eligible.argtypes = [P]
    lib.molann_b200_prepared_bytes.restype = sz
    lib.molann_b200_prepared_bytes.argtypes = [P]
    lib.molann_b200_prepare.argtypes = [P, vp, sz, vp, ctypes.POINTER(vp)]
    lib.molann_b200_prepared_refresh.argtypes = [vp, P, vp]
    lib.molann_b200_prepared_workspace_bytes.restype = sz
    lib.molann_b200_prepared_workspace_bytes.argtypes = [vp, i64]
    lib.molann_b200_forward_prepared.argtypes = [vp, P, vp, i64, vp, vp, sz, vp]
    lib.molann_b200_value_and_grad_prepared.argtypes = [vp, P, vp, vp, i64, vp, vp, vp, sz, vp]
    lib.molann_b200_prepared_destroy.restype = None
    lib.molann_b200_prepared_destroy.argtypes = [vp]
    D = ctypes.POINTER(MolannDecoder)
    lib.molann_b200_train_eligible.argtypes = [P, D]
    lib.molann_b200_train_param_count.restype = sz
    lib.molann_b200_train_param_count.argtypes = [P, D]
    lib.molann_b200_train_workspace_bytes.restype = sz
    lib.molann_b200_train_workspace_bytes.argtypes = [P, D]
    lib.molann_b200_train_loss_and_grads.argtypes = [P, D, vp, i64, ctypes.c_float, vp, vp, sz, vp]
    lib.molann_b200_sgd_apply.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.c_int32, vp, ctypes.c_float,
                                          vp]
    lib.molann_b200_allreduce_buffer_bytes.restype = sz
    lib.molann_b200_allreduce_buffer_bytes.argtypes = [i64, ctypes.c_int32]
    lib.molann_b200_allreduce_sgd.argtypes = [vp, vp, i64, ctypes.POINTER(vp), ctypes.c_int32, ctypes.c_int32, vp,
                                              ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.c_int32, ctypes.c_float, vp]
    _cabi = lib
    return lib


def load_torch_ops():
    global _ops_loaded
    if _ops_loaded:
        return
    if not os.path.isfile(LIB_TORCH):
        raise RuntimeError(_HELP % (LIB_TORCH, "file not found"))
    try:
        torch.ops.load_library(LIB_TORCH)
    except OSError as exc:
        raise RuntimeError(_HELP % (LIB_TORCH, exc)) from exc
    _ops_loaded = True


def check(status, what="call"):
    if status != 0:
        lib = cabi()
        msg = lib.molann_b200_strerror(status).decode()
        if status == 5:
            code = lib.molann_b200_last_cuda_error()
            msg += " [cudaError %d: %s]" % (code, lib.molann_b200_cuda_error_string(code).decode())
        raise RuntimeError("molann_b200 %s failed: %s" % (what, msg))


def launch_count():
    return int(cabi().molann_b200_launch_count())
