"""Jacobian / all-forces entry point and the libtorch-only consumer (-m gpu), SURVEY 8(f) item 1.

`MolANN.value_and_jacobian(x)` returns y and J[o] = d y[:, o] / dx for every output o from ONE pass over x -- what
an MD plugin that loads the exported TorchScript model (reference README.rst:51, molann/ann.py:109-111) needs per
step.  The oracle is the reference computation: one autograd call per output on the fp64 restatement."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT, S, assert_parity, oracle_model

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in list(os.environ):
        if k.startswith("MOLANN_B200_"):
            monkeypatch.delenv(k, raising=False)
    yield


def _weights(model, nl):
    sd = model.state_dict()
    return ([sd["ann_layers.%dth_layer.weight" % (k + 1)] for k in range(nl)],
            [sd["ann_layers.%dth_layer.bias" % (k + 1)] for k in range(nl)])


def _oracle_jacobian(spec, ws, bs, x, dtype):
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    y = oracle_model(spec, ws, bs, dtype)(xx)
    planes = [torch.autograd.grad(y[:, o].sum(), xx, retain_graph=True)[0] for o in range(y.shape[1])]
    return y.detach(), torch.stack(planes)


@pytest.mark.parametrize("variant", ["C2", "one_hidden", "three_outputs", "mixed_program", "no_alignment", "relu",
                                     "C1", "C3s"])
def test_value_and_jacobian_against_oracle(variant):
    """C2 class (ONE fused launch) incl. one hidden layer, 3 and 8-wide outputs, mixed programs, no alignment; C1 (FFMA
    family) and C3s take one value-and-gradient pass per output through the same entry point."""
    from molann_b200 import _lib
    name = variant if variant in ("C1", "C3s") else "C2"
    spec = S.get_spec(name)
    if variant == "one_hidden":
        spec.layer_dims = [30, 48, 2]
    elif variant == "three_outputs":
        spec.layer_dims = [30, 64, 64, 3]
    elif variant == "mixed_program":
        spec.features = [("p", "position", [1, 4, 5, 6, 8]), ("d", "dihedral", [4, 6, 8, 14]), ("b", "bond", [8, 10]),
                         ("a", "angle", [6, 8, 14]), ("p2", "position", [16, 18])]
        spec.layer_dims = [spec.feature_dim(), 32, 40, 2]
    elif variant == "no_alignment":
        spec.align_ix = None
        spec.trans_sigma, spec.rotate = 0.0, False
    elif variant == "relu":
        spec.activation = "relu"
    model, _ = S.build_model(spec, init_seed=4)
    nl = len(spec.layer_dims) - 1
    ws, bs = _weights(model, nl)
    L = 517                                            # four full tiles and a ragged one
    x = S.make_frames(spec, L, seed=21)
    y64, j64 = _oracle_jacobian(spec, ws, bs, x, torch.float64)
    y32, j32 = _oracle_jacobian(spec, ws, bs, x, torch.float32)
    model = model.cuda()
    before = _lib.launch_count()
    y, jac = model.value_and_jacobian(x.cuda())
    torch.cuda.synchronize()
    launches = _lib.launch_count() - before
    k = spec.out_dim()
    assert tuple(jac.shape) == (k, L, spec.n_inp, 3) and tuple(y.shape) == (L, k)
    assert not y.requires_grad and not jac.requires_grad
    if name == "C2":
        assert launches == 1, launches                 # every plane while the tile is on chip
    assert_parity(y.cpu(), y64, y32, TOL, variant + " y")
    for o in range(k):
        if variant == "relu":
            from helpers import frame_rel_err
            assert float((frame_rel_err(jac[o].cpu(), j64[o]) > TOL).float().mean()) < 0.02
        else:
            # a single output's gradient can nearly vanish on a frame (C1: a 5-unit tanh net on two features); measure
            # against at least half the plane's typical magnitude there
            fl = 0.5 * float(j64[o].abs().amax(dim=(1, 2)).median())
            assert_parity(jac[o].cpu(), j64[o], j32[o], TOL, "%s J[%d]" % (variant, o), floor=fl)
    # consistency with the one-cotangent entry point: <cot, J> == value_and_grad(cot)
    cot = torch.randn(L, k, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    _, gx = model.value_and_grad(x.cuda(), cot)
    comb = (jac * cot.t().reshape(k, L, 1, 1)).sum(0)
    assert float((comb - gx).abs().max()) <= 2e-5 * float(gx.abs().max())


def test_jacobian_small_batches_and_unaligned():
    """L = 1 ... 130 (what an MD plugin sends), and an input 4 bytes off the 16-byte grid (non-TMA path)."""
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    ws, bs = _weights(model, 3)
    model = model.cuda()
    for L in (1, 2, 3, 127, 128, 130):
        x = S.make_frames(spec, L, seed=50 + L)
        y64, j64 = _oracle_jacobian(spec, ws, bs, x, torch.float64)
        y32, j32 = _oracle_jacobian(spec, ws, bs, x, torch.float32)
        y, jac = model.value_and_jacobian(x.cuda())
        assert_parity(y.cpu(), y64, y32, TOL, "L=%d y" % L)
        for o in range(2):
            assert_parity(jac[o].cpu(), j64[o], j32[o], TOL, "L=%d J[%d]" % (L, o))
        flat = torch.zeros(x.numel() + 1, device="cuda")
        xv = flat[1:].view(x.shape)
        xv.copy_(x)
        y2, jac2 = model.value_and_jacobian(xv)
        assert torch.equal(y2, y) and torch.equal(jac2, jac)


def test_libtorch_consumer(tmp_path):
    """A C++ program that links only libtorch: dlopen()s the shim, loads the exported model, runs forward + C++ autograd
    and the exported value_and_grad / value_and_jacobian methods, and must reproduce what Python computed."""
    from molann_b200 import build
    consumer = build.build_consumer()
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    L = 300
    x = S.make_frames(spec, L, device="cuda", seed=3)
    cot = torch.randn(L, 2, device="cuda")
    xg = x.clone().requires_grad_(True)
    y = model(xg)
    (gx,) = torch.autograd.grad(y, xg, cot)
    _, jac = model.value_and_jacobian(x)
    scripted = torch.jit.script(model)
    mpath = os.path.join(tmp_path, "model.pt")
    scripted.save(mpath)

    class IO(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for name, t in (("x", x), ("cot", cot), ("y", y.detach()), ("gx", gx), ("jac", jac)):
                self.register_buffer(name, t.detach().cpu())

        def forward(self):
            return self.x

    ipath = os.path.join(tmp_path, "io.pt")
    torch.jit.script(IO()).save(ipath)
    shim = os.path.join(ROOT, "molann_b200", "libmolann_b200_torch.so")
    out = subprocess.run([consumer, shim, mpath, ipath], capture_output=True, text=True, timeout=600)
    assert "CONSUMER_OK" in out.stdout, out.stdout + out.stderr
