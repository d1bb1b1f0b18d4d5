"""The C-ABI shared library: loads without a GPU, exports every symbol include/molann_b200.h declares,
validates plans on the host, and refuses to compute when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT, S, golden, golden_weights, spec_program
from molann_b200 import _lib


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "molann_b200.h")).read()
    return sorted(set(re.findall(r"\b(molann_b200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.cabi()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in include/molann_b200.h is not exported" % s
    assert lib.molann_b200_version() == 100
    assert lib.molann_b200_strerror(0) == b"ok" and b"plan" in lib.molann_b200_strerror(2)


def test_struct_layout_matches_header():
    # 2 int32, 2 ptr, int32 (+pad), ptr, 4 int32, 9 int32 (+pad), 8 ptr, 8 ptr
    assert ctypes.sizeof(_lib.MolannPlan) == 8 + 16 + 8 + 8 + 16 + 36 + 4 + 64 + 64


def host_plan(spec, n_layers=None):
    """Plan with fake (non-null) device pointers: only scalar validation is exercised."""
    aidx, ref, feats, entries, d = spec_program(spec)
    p = _lib.MolannPlan()
    p.n_inp = spec.n_inp
    if aidx is not None:
        p.n_align, p.align_idx, p.ref_x = len(aidx), 0x1000, 0x2000
    p.n_entries, p.entries, p.d_feat, p.use_angle_value = entries.shape[0], 0x3000, d, int(spec.use_angle_value)
    dims = spec.layer_dims
    p.n_layers = len(dims) - 1 if n_layers is None else n_layers
    for k, v in enumerate(dims):
        p.dims[k] = v
    for k in range(len(dims) - 1):
        p.W[k], p.b[k] = 0x4000 + k * 0x100, 0x8000 + k * 0x100
    return p


def test_plan_validate_statuses():
    lib = _lib.cabi()
    p = host_plan(S.get_spec("C2"))
    assert lib.molann_b200_plan_validate(ctypes.byref(p)) == 0
    assert lib.molann_b200_plan_validate(None) == 1
    bad = host_plan(S.get_spec("C2")); bad.dims[0] = 31
    assert lib.molann_b200_plan_validate(ctypes.byref(bad)) == 2
    bad = host_plan(S.get_spec("C2")); bad.W[1] = None
    assert lib.molann_b200_plan_validate(ctypes.byref(bad)) == 1
    bad = host_plan(S.get_spec("C2")); bad.act_id = 9
    assert lib.molann_b200_plan_validate(ctypes.byref(bad)) == 2
    bad = host_plan(S.get_spec("C2")); bad.n_layers = 9
    assert lib.molann_b200_plan_validate(ctypes.byref(bad)) == 2
    bad = host_plan(S.get_spec("C2")); bad.entries = None
    assert lib.molann_b200_plan_validate(ctypes.byref(bad)) == 1
    bad = host_plan(S.get_spec("C2")); bad.n_inp = 0
    assert lib.molann_b200_plan_validate(ctypes.byref(bad)) == 2


def test_dispatch_table_and_workspace():
    """Small systems go to the fused kernel, big ones to the general path; workspace is O(chunk)."""
    lib = _lib.cabi()
    for name, want in (("C1", 1), ("C2", 1), ("C3", 0), ("C5", 0)):
        p = host_plan(S.get_spec(name))
        assert lib.molann_b200_path_for(ctypes.byref(p), 0) == want, name
        assert lib.molann_b200_path_for(ctypes.byref(p), 1) == want, name
    p = host_plan(S.get_spec("C3"))
    w1 = lib.molann_b200_workspace_bytes(ctypes.byref(p), 1 << 20, 0)
    w2 = lib.molann_b200_workspace_bytes(ctypes.byref(p), 1 << 24, 0)
    assert w1 == w2 > 0
    assert lib.molann_b200_workspace_bytes(ctypes.byref(p), 1 << 20, 1) > w1
    assert lib.molann_b200_workspace_bytes(ctypes.byref(p), 100, 0) < w1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_wide_path_eligibility(monkeypatch):
    """Big systems with a wide first layer take the prepared (fused wide) forward; small ones keep their kernels."""
    lib = _lib.cabi()
    for k in list(os.environ):
        if k.startswith("MOLANN_B200_"):
            monkeypatch.delenv(k, raising=False)
    for name, want in (("C1", 0), ("C2", 0), ("C3", 1), ("C5", 1)):
        p = host_plan(S.get_spec(name))
        assert lib.molann_b200_wide_eligible(ctypes.byref(p)) == want, name
    p = host_plan(S.get_spec("C3"))
    nbytes = lib.molann_b200_prepared_bytes(ctypes.byref(p))
    assert 800 * 256 * 8 < nbytes < 64 << 20              # hi + lo copies of W1 and some tables
    monkeypatch.setenv("MOLANN_B200_WIDE", "0")
    assert lib.molann_b200_wide_eligible(ctypes.byref(p)) == 0
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    assert lib.molann_b200_wide_eligible(ctypes.byref(host_plan(S.get_spec("C2")))) == 1
    four = host_plan(S.get_spec("C2"), n_layers=None)
    four.n_layers = 4
    four.dims[4] = 2
    four.W[3], four.b[3] = 0x9000, 0x9100
    assert lib.molann_b200_wide_eligible(ctypes.byref(four)) == 0     # three hidden layers: not this kernel
    assert lib.molann_b200_prepared_workspace_bytes(None, 100) == 0


def test_compute_entry_points_fail_without_gpu():
    lib = _lib.cabi()
    p = host_plan(S.get_spec("C2"))
    st = lib.molann_b200_forward(ctypes.byref(p), 0x10000, 8, 0x20000, None, 0, None)
    assert st == 5                                     # MOLANN_ERR_CUDA, never a silent CPU result
    assert lib.molann_b200_forward(ctypes.byref(p), None, 8, 0x20000, None, 0, None) == 1
    assert lib.molann_b200_forward(ctypes.byref(p), 0x10000, 0, 0x20000, None, 0, None) == 0     # L == 0: no-op
    assert lib.molann_b200_forward(ctypes.byref(p), 0x10001, 8, 0x20000, None, 0, None) == 4
    q = host_plan(S.get_spec("C2"), n_layers=0)
    assert lib.molann_b200_forward(ctypes.byref(q), 0x10000, 8, 0x20000, None, 0, None) == 6


def host_decoder(dims, act_id=0):
    d = _lib.MolannDecoder()
    d.n_layers, d.act_id = len(dims) - 1, act_id
    for k, v in enumerate(dims):
        d.dims[k] = v
    for k in range(len(dims) - 1):
        d.W[k], d.b[k] = 0xA000 + k * 0x100, 0xB000 + k * 0x100
    return d


def test_training_entry_points_validate_on_host():
    """molann_b200_train_* (include/molann_b200.h): scalar validation and the flat layout need no GPU; compute fails
    with MOLANN_ERR_CUDA here, never with a CPU result."""
    lib = _lib.cabi()
    assert ctypes.sizeof(_lib.MolannDecoder) == 8 + 36 + 4 + 64 + 64
    enc = host_plan(S.get_spec("C2"))
    dec = host_decoder([2, 64, 64, 30])
    P = (30 + 1) * 64 + (64 + 1) * 64 + (64 + 1) * 2 + (2 + 1) * 64 + (64 + 1) * 64 + (64 + 1) * 30
    assert lib.molann_b200_train_param_count(ctypes.byref(enc), ctypes.byref(dec)) == P == 12576
    assert lib.molann_b200_train_param_count(ctypes.byref(enc), None) == 0
    assert lib.molann_b200_train_param_count(ctypes.byref(enc), ctypes.byref(host_decoder([3, 64, 30]))) == 0
    assert lib.molann_b200_train_param_count(ctypes.byref(enc), ctypes.byref(host_decoder([2, 64, 29]))) == 0
    assert lib.molann_b200_train_eligible(ctypes.byref(enc), ctypes.byref(dec)) == 0        # no device here
    a = (0x10000, 128, ctypes.c_float(1.0), 0x20000, 0x30000, 1 << 30, None)
    assert lib.molann_b200_train_loss_and_grads(ctypes.byref(enc), ctypes.byref(dec), *a) == 5
    assert lib.molann_b200_train_loss_and_grads(ctypes.byref(enc), None, *a) == 1
    assert lib.molann_b200_train_loss_and_grads(ctypes.byref(enc), ctypes.byref(host_decoder([3, 64, 30])), *a) == 2
    assert lib.molann_b200_train_loss_and_grads(ctypes.byref(enc), ctypes.byref(dec), 0x10000, 128,
                                                ctypes.c_float(1.0), None, 0x30000, 1 << 30, None) == 1
    assert lib.molann_b200_train_loss_and_grads(ctypes.byref(enc), ctypes.byref(dec), 0x10000, 128,
                                                ctypes.c_float(1.0), 0x20002, 0x30000, 1 << 30, None) == 4
    ptrs = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    numel = (ctypes.c_int64 * 2)(4, 4)
    assert lib.molann_b200_sgd_apply(ptrs, numel, 0, 0x3000, ctypes.c_float(0.1), None) == 0     # nothing to do
    assert lib.molann_b200_sgd_apply(None, numel, 2, 0x3000, ctypes.c_float(0.1), None) == 1
    assert lib.molann_b200_sgd_apply(ptrs, numel, 99, 0x3000, ctypes.c_float(0.1), None) == 2
    assert lib.molann_b200_sgd_apply(ptrs, numel, 2, 0x3000, ctypes.c_float(0.1), None) == 5


def test_allreduce_sgd_entry_point_validates_on_host():
    """molann_b200_allreduce_sgd (include/molann_b200.h): buffer sizing and argument checks need no GPU."""
    lib = _lib.cabi()
    assert lib.molann_b200_allreduce_buffer_bytes(12577, 8) == (2 * 12577 * 4 + 255) // 256 * 256 + 256
    assert lib.molann_b200_allreduce_buffer_bytes(12577, 9) == 0 and lib.molann_b200_allreduce_buffer_bytes(0, 2) == 0
    bufs = (ctypes.c_void_p * 2)(0x100000, 0x200000)
    ptrs = (ctypes.c_void_p * 1)(0x1000)
    numel = (ctypes.c_int64 * 1)(16)
    call = lib.molann_b200_allreduce_sgd
    assert call(0x4000, 0x5000, 17, bufs, 0, 2, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 5    # no device
    assert call(None, 0x5000, 17, bufs, 0, 2, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 1
    assert call(0x4000, 0x5000, 17, bufs, 2, 2, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 2    # rank >= world
    assert call(0x4000, 0x5000, 17, bufs, 0, 9, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 2
    assert call(0x4000, 0x5000, 8, bufs, 0, 2, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 2     # params > vector
    assert call(0x4002, 0x5000, 17, bufs, 0, 2, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 4
    bad = (ctypes.c_void_p * 2)(0x100000, None)
    assert call(0x4000, 0x5000, 17, bad, 0, 2, 0x6000, ptrs, numel, 1, ctypes.c_float(0.1), None) == 1
