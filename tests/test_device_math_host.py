"""The CUDA device functions of molann_b200/csrc/geometry.cuh compiled for the host (tests/host/) and
checked against the oracle: validates Jacobi/quaternion alignment, the feature program and the
closed-form backward without a GPU.  (The real kernels are checked by the -m gpu suites.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers import (ROOT, S, golden, oracle_preprocess, oracle_value_and_grad, spec_program, frame_rel_err, R)

HOST_DIR = os.path.join(ROOT, "tests", "host")
LIB = os.path.join(HOST_DIR, "libmolann_host_emu.so")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HOST_DIR, "host_emulation.cpp")
    hdr = os.path.join(ROOT, "molann_b200", "csrc", "geometry.cuh")
    if not os.path.isfile(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB, src], check=True)
    return ctypes.CDLL(LIB)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def emu_run(emu, spec, x, gfeat):
    aidx, ref, feats, entries, d = spec_program(spec)
    a = np.asarray(aidx if aidx is not None else [], dtype=np.int32)
    r = np.ascontiguousarray(ref.numpy(), dtype=np.float32)
    L, n = x.shape[0], x.shape[1]
    feat = np.zeros((L, d), np.float32)
    gx = np.zeros_like(x)
    emu.emu_preprocess_forward(n, len(a), ptr(a), ptr(r), len(entries), ptr(entries), d, int(spec.use_angle_value),
                               ptr(x), ctypes.c_longlong(L), ptr(feat))
    emu.emu_preprocess_backward(n, len(a), ptr(a), ptr(r), len(entries), ptr(entries), d, int(spec.use_angle_value),
                                ptr(x), ptr(gfeat), ctypes.c_longlong(L), ptr(gx))
    return feat, gx


@pytest.mark.parametrize("name,L", [("C1", 300), ("C2", 600), ("C3s", 64)])
def test_device_math_vs_oracle(emu, name, L):
    spec = S.get_spec(name)
    x = S.make_frames(spec, L, seed=4242).numpy()
    gfeat = np.random.RandomState(3).randn(L, spec.feature_dim()).astype(np.float32)
    feat, gx = emu_run(emu, spec, x, gfeat)
    f64, gx64 = oracle_value_and_grad(oracle_preprocess(spec), torch.from_numpy(x), torch.from_numpy(gfeat))
    assert float(frame_rel_err(feat, f64).max()) < 1e-5
    assert float(frame_rel_err(gx, gx64).max()) < 1e-5


@pytest.mark.parametrize("name", ["C1", "C2", "C3s"])
def test_device_math_vs_reference_goldens(emu, name):
    spec = S.get_spec(name)
    g = golden("config_" + name)
    feat, gx = emu_run(emu, spec, np.ascontiguousarray(g["x"]), np.ascontiguousarray(g["cotf"]))
    assert float(frame_rel_err(feat, g["feat64"]).max()) < 1e-5
    assert float(frame_rel_err(gx, g["gxf64"]).max()) < 1e-5


def test_alignment_standalone_and_reflection_frames(emu):
    """Stand-alone alignment incl. frames whose SVD route needs the det<0 correction, n_a = 3."""
    rng = np.random.RandomState(11)
    n, L = 9, 400
    for aidx in ([0, 3, 7], [1, 2, 4, 5, 8]):
        ref = rng.randn(len(aidx), 3).astype(np.float32)
        ref -= ref.mean(0)
        x = (rng.randn(L, n, 3) * 2 + 30).astype(np.float32)      # random clouds: ~half need the correction
        a = np.asarray(aidx, np.int32)
        out = np.zeros_like(x)
        emu.emu_align_forward(n, len(a), ptr(a), ptr(ref), ptr(x), ctypes.c_longlong(L), ptr(out))
        xt = torch.from_numpy(x).double().requires_grad_(True)
        z = R.align_forward(xt, aidx, torch.from_numpy(ref).double())
        u, s, vh = torch.linalg.svd(torch.matmul((xt[:, aidx] - xt[:, aidx].mean(1, True)).permute(0, 2, 1).detach(),
                                                 torch.from_numpy(ref).double()))
        n_reflect = int((torch.linalg.det(u @ vh) < 0).sum())
        assert n_reflect > L // 10
        gap = (s[:, 1] + s[:, 2] * torch.sign(torch.linalg.det(u @ vh))) / s[:, 0]      # conditioning
        ok = gap.abs() > 0.05
        assert float(frame_rel_err(out, z.detach())[ok].max()) < 2e-5
        gout = rng.randn(L, n, 3).astype(np.float32)
        (gx64,) = torch.autograd.grad((z * torch.from_numpy(gout).double()).sum(), xt)
        gx = np.zeros_like(x)
        emu.emu_align_backward(n, len(a), ptr(a), ptr(ref), ptr(x), ptr(gout), ctypes.c_longlong(L), ptr(gx))
        assert float(frame_rel_err(gx, gx64)[ok].max()) < 5e-4
        assert float(frame_rel_err(gx, gx64)[ok].median()) < 1e-5


def _svd_rotation64(x, aidx, ref):
    xs = x[:, aidx, :].astype(np.float64)
    xt = xs - xs.mean(1, keepdims=True)
    H = np.einsum("nka,kb->nab", xt, ref.astype(np.float64))
    U, s, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(U @ Vt))
    D = np.zeros_like(H)
    D[:, 0, 0] = 1
    D[:, 1, 1] = 1
    D[:, 2, 2] = d
    return U @ D @ Vt, (s[:, 1] + s[:, 2] * d) / s[:, 0]


def test_fast_rotation_path_accuracy_and_fallback(emu):
    """The polynomial (QCP) + first-order-correction path must (a) be the one taken on well-conditioned
    trajectories, (b) agree with the fp64 SVD route of the reference (ann.py:188-195) to fp32 rounding, and
    (c) hand near-degenerate frames to the Jacobi fallback instead of returning a poor rotation."""
    spec = S.get_spec("C2")
    aidx, ref, _, _, _ = spec_program(spec)
    a = np.asarray(aidx, np.int32)
    r = np.ascontiguousarray(ref.numpy(), np.float32)
    L = 20000
    x = S.make_frames(spec, L, seed=99).numpy()
    Rk = np.zeros((L, 9), np.float32)
    ok = np.zeros(L, np.int32)
    emu.emu_kabsch(x.shape[1], len(a), ptr(a), ptr(r), ptr(x), ctypes.c_longlong(L), ptr(Rk), ptr(ok))
    R64, gap = _svd_rotation64(x, aidx, r)
    assert ok.mean() > 0.999
    assert np.abs(Rk.reshape(L, 3, 3) - R64).max() < 1.5e-6
    # random point clouds: many reflections / small gaps -> fallback must keep the answer right
    rng = np.random.RandomState(5)
    x2 = (rng.randn(L, x.shape[1], 3) * 2 + 10).astype(np.float32)
    emu.emu_kabsch(x.shape[1], len(a), ptr(a), ptr(r), ptr(x2), ctypes.c_longlong(L), ptr(Rk), ptr(ok))
    R64, gap = _svd_rotation64(x2, aidx, r)
    err = np.abs(Rk.reshape(L, 3, 3) - R64).max((1, 2))
    assert 0.5 < ok.mean() < 1.0                       # both paths exercised
    cond = np.abs(gap) > 0.05
    assert err[cond].max() < 2e-5
    assert err[cond & (ok == 1)].max() < 2e-5
    assert np.isfinite(Rk).all()


def test_activations(emu):
    emu.emu_act_forward.restype = ctypes.c_float
    emu.emu_act_forward.argtypes = [ctypes.c_float, ctypes.c_int]
    emu.emu_act_grad.restype = ctypes.c_float
    emu.emu_act_grad.argtypes = [ctypes.c_float, ctypes.c_int]
    for v in (-3.0, -0.2, 0.0, 0.7, 4.0):
        t = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        for act_id, fn in ((0, torch.tanh), (1, torch.relu), (2, torch.sigmoid)):
            h = fn(t)
            (g,) = torch.autograd.grad(h, t)
            assert abs(emu.emu_act_forward(v, act_id) - float(h)) < 1e-6
            assert abs(emu.emu_act_grad(float(h), act_id) - float(g)) < 1e-6
