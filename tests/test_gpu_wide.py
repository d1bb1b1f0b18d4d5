"""The fused wide kernel (csrc/fused_wide.cuh, `molann_b200_forward_prepared`) against the oracle (-m gpu).

It is the forward of MolANN (reference molann/ann.py:620-624 over :553-565) for big systems with a wide first layer
(BASELINE configs[2], [4]): ONE persistent kernel per call.  The small cases below force it onto plans the small-system
kernels would normally take (MOLANN_B200_WIDE=1) so that the oracle finishes in seconds; the full-width cases use it by
default."""
import ctypes
import os

import numpy as np
import pytest
import torch

from helpers import (S, CPlan, assert_parity, golden, golden_weights, oracle_model, oracle_value_and_grad)

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in list(os.environ):
        if k.startswith("MOLANN_B200_"):
            monkeypatch.delenv(k, raising=False)
    yield


def dev(t):
    return torch.as_tensor(t).cuda().contiguous()


def _weights(model, nl):
    sd = model.state_dict()
    return ([sd["ann_layers.%dth_layer.weight" % (k + 1)] for k in range(nl)],
            [sd["ann_layers.%dth_layer.bias" % (k + 1)] for k in range(nl)])


def _oracle(spec, ws, bs, x):
    y64 = oracle_model(spec, ws, bs)(x)
    y32 = oracle_model(spec, ws, bs, torch.float32)(x)
    return y64, y32


@pytest.mark.parametrize("name", ["C2", "C3s"])
def test_prepared_forward_against_reference_goldens(name, monkeypatch):
    """C ABI: prepare once, forward on the reference's golden inputs (C2: positions only; C3s: positions + dihedrals)."""
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    spec = S.get_spec(name)
    g = golden("config_" + name)
    ws, bs = golden_weights(g, len(spec.layer_dims) - 1)
    plan = CPlan(spec, ws, bs)
    assert plan.lib.molann_b200_wide_eligible(ctypes.byref(plan.p)) == 1
    y = plan.forward_prepared(dev(g["x"]))
    assert_parity(y.cpu(), g["y64"], g["y32"], TOL, "%s y (prepared)" % name)
    y2 = plan.forward_prepared(dev(g["x"]))
    assert torch.equal(y, y2)                                     # deterministic: fixed summation order


VARIANTS = ["one_hidden", "three_outputs", "narrow_wide", "mixed_program", "no_alignment", "sigmoid_mixed", "relu",
            "many_invariants", "angle_values"]


@pytest.mark.parametrize("ku", ["4", "2", "4-global-tables"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_wide_kernel_variants(variant, ku, monkeypatch):
    """Shapes and programs: one hidden layer, widths that are not multiples of 64, up to 8 outputs, programs mixing all
    feature types in any order, more invariant columns than position atoms (units of four invariants), none at all,
    no alignment, the other activations, angle values -- on both K-chunk sizes of the kernel (KU = 4: K = 16 per operand
    stage, the default; KU = 2: the half-size stages that 60 KB frames need)."""
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    monkeypatch.setenv("MOLANN_B200_WIDE_KU", ku[0])
    if ku.endswith("global-tables"):                               # plan tables read from global memory (the instantiation
        monkeypatch.setenv("MOLANN_B200_WIDE_TABLES", "0")         # that plans too big for shared memory take)
    spec = S.get_spec("C2")
    if variant == "one_hidden":
        spec.layer_dims = [30, 48, 2]
    elif variant == "three_outputs":
        spec.layer_dims = [30, 64, 64, 3]
    elif variant == "narrow_wide":
        spec.layer_dims = [30, 16, 80, 8]
    elif variant in ("mixed_program", "sigmoid_mixed", "angle_values"):
        spec.features = [("d", "dihedral", [4, 6, 8, 14]), ("p", "position", [1, 4, 5, 6, 8]), ("b", "bond", [8, 10]),
                         ("a", "angle", [6, 8, 14]), ("p2", "position", [16, 18])]
        if variant == "angle_values":
            spec.use_angle_value = True
        spec.layer_dims = [spec.feature_dim(), 32, 40, 2]
        if variant == "sigmoid_mixed":
            spec.activation = "sigmoid"
    elif variant == "no_alignment":
        spec.align_ix = None
        spec.trans_sigma, spec.rotate = 0.0, False
    elif variant == "relu":
        spec.activation = "relu"
    elif variant == "many_invariants":
        spec.features = [("p", "position", [1, 4])]
        for k in range(9):
            spec.features.append(("d%d" % k, "dihedral", [k, k + 1, k + 2, k + 4]))
            spec.features.append(("b%d" % k, "bond", [k, k + 3]))
        spec.layer_dims = [spec.feature_dim(), 96, 32, 2]
    model, _ = S.build_model(spec, init_seed=3)
    nl = len(spec.layer_dims) - 1
    ws, bs = _weights(model, nl)
    L = 777
    x = S.make_frames(spec, L, seed=12)
    y64, y32 = _oracle(spec, ws, bs, x)
    model = model.cuda()
    from molann_b200 import _lib
    with torch.no_grad():
        y = model(dev(x))                                          # builds the prepared plan (pack launches)
        before = _lib.launch_count()
        y_again = model(dev(x))
    assert _lib.launch_count() == before + 1                       # steady state: ONE kernel, no packing
    assert torch.equal(y, y_again)
    if variant == "relu":                                          # a unit within fp32 noise of its kink may flip
        from helpers import frame_rel_err
        assert float((frame_rel_err(y.cpu(), y64) > TOL).float().mean()) < 0.02
    else:
        assert_parity(y.cpu(), y64, y32, TOL, variant + " y")


@pytest.mark.parametrize("L", [1, 2, 5, 31, 32, 33, 127, 128, 129, 255, 257, 1000])
def test_wide_ragged_frame_counts_and_unaligned_input(L, monkeypatch):
    """Partial tiles and sub-tiles; x 4 / 8 / 12 bytes off the 16-byte grid (the frame ring copies from the boundary
    below; the batch's last frame then takes the plain-load path)."""
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    spec = S.get_spec("C3s")
    g = golden("config_C3s")
    ws, bs = golden_weights(g, 3)
    plan = CPlan(spec, ws, bs)
    x = S.make_frames(spec, L, seed=900 + L)
    y64, y32 = _oracle(spec, ws, bs, x)
    y = plan.forward_prepared(dev(x))
    assert_parity(y.cpu(), y64, y32, TOL, "C3s L=%d y" % L)
    flat = torch.zeros(x.numel() + 3, device="cuda")
    for shift in (1, 2, 3):
        xv = flat[shift:shift + x.numel()].view(x.shape)
        xv.copy_(x)
        assert xv.data_ptr() % 16 == 4 * shift
        assert torch.equal(plan.forward_prepared(xv), y), shift


def test_wide_slices_are_bitwise_identical(monkeypatch):
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    spec = S.get_spec("C3s")
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, 3000, device="cuda", seed=77)
    with torch.no_grad():
        y = model(x)
        for a, n in ((1, 128), (3, 1000), (2, 129), (127, 1), (5, 2995)):
            assert torch.equal(model(x[a:a + n]), y[a:a + n]), (a, n)


def test_wide_weight_update_repacks(monkeypatch):
    """Training changes the weights in place: the cached prepared plan must follow (`_version`)."""
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    spec = S.get_spec("C3s")
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, 300, device="cuda", seed=3)
    with torch.no_grad():
        y0 = model(x)
        for p in model.parameters():
            p.mul_(0.5)
        y1 = model(x)
    ws, bs = _weights(model, 3)
    y64, y32 = _oracle(spec, [w.cpu() for w in ws], [b.cpu() for b in bs], x.cpu())
    assert not torch.allclose(y0, y1)
    assert_parity(y1.cpu(), y64, y32, TOL, "after in-place weight update")


@pytest.mark.parametrize("name,L", [("C3", 203), ("C5", 131)])
def test_wide_full_width_default_path(name, L):
    """BASELINE configs[2] / [4] shapes: the fused wide kernel IS the default forward, one launch per call (C5's 60 KB
    frames take the K = 8 operand stages so that two frames fit next to them in shared memory), and the result must
    match the oracle."""
    from molann_b200 import _lib
    spec = S.get_spec(name)
    model, _ = S.build_model(spec)
    ws, bs = _weights(model, 3)
    x = S.make_frames(spec, L, seed=17)
    y64, y32 = _oracle(spec, ws, bs, x)
    model = model.cuda()
    with torch.no_grad():
        y = model(dev(x))
        before = _lib.launch_count()
        y2 = model(dev(x))
    n_launch = _lib.launch_count() - before
    assert n_launch == 1
    assert torch.equal(y, y2)
    assert_parity(y.cpu(), y64, y32, TOL, name + " y (fused wide kernel)")
    # value-and-gradient of a prepared wide plan: the SAME fused forward kernel (it also leaves the hidden activations
    # behind), then narrow + 2 GEMMs backward-to-input on the operands packed once + the block preprocess backward
    # = 5 launches, no forward recompute, no packing; y and gx match the fp64 oracle
    cot = torch.randn(L, 2, generator=torch.Generator().manual_seed(5))
    y64b, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
    _, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
    before = _lib.launch_count()
    yl, gx = model.value_and_grad(dev(x), cot.cuda())
    assert _lib.launch_count() - before == 5
    # (the activation-storing instantiation of the kernel is a different compilation: equal to rounding, not bitwise)
    assert float((yl - y).abs().max()) <= 2e-6 * max(1.0, float(y.abs().max()))
    assert_parity(yl.cpu(), y64b, None, TOL, name + " y (value_and_grad)")
    assert_parity(gx.cpu(), gx64, gx32, TOL, name + " gx (fused forward + layered backward)")


@pytest.mark.parametrize("name", ["C3", "C5"])
def test_wide_bench_size(name):
    """Bench-size batch (32768 x 2000 atoms / 16384 x 5000 atoms): finite, slices bitwise identical, rigid-motion
    invariance, and a 48-frame oracle sample."""
    spec = S.get_spec(name)
    L = spec.default_frames
    model, _ = S.build_model(spec)
    ws, bs = _weights(model, 3)
    model = model.cuda()
    x = S.make_frames(spec, L, device="cuda", seed=9)
    from molann_b200 import _lib
    with torch.no_grad():
        y = model(x)
        assert torch.isfinite(y).all()
        before = _lib.launch_count()
        ys = model(x[1000:1333].contiguous())
        if _lib.launch_count() - before == 1:                       # fused wide kernel: frames bitwise independent
            assert torch.equal(ys, y[1000:1333])
        else:                                                        # layered fallback: K order follows the tile's CTA
            assert float((ys - y[1000:1333]).abs().max()) <= 2e-6 * max(1.0, float(y.abs().max()))
        Ls = 2048
        Rm = S.random_rotations(Ls, torch.Generator(device="cuda").manual_seed(7), "cuda")
        y2 = model((torch.bmm(x[:Ls], Rm) + 5.0).contiguous())
        assert float((y2 - y[:Ls]).abs().max()) < 3e-4 * max(1.0, float(y[:Ls].abs().max()))
    idx = torch.randint(0, L, (48,), generator=torch.Generator().manual_seed(0))
    xs = x[idx.cuda()].cpu()
    y64, y32 = _oracle(spec, ws, bs, xs)
    assert_parity(y[idx.cuda()].cpu(), y64, y32, TOL, name + " bench-size y sample")


@pytest.mark.parametrize("activation", ["tanh", "sigmoid"])
def test_wide_value_and_grad_chunks_and_activations(activation, monkeypatch):
    """value_and_grad of a prepared wide plan: (a) cut into several chunks (MOLANN_B200_WIDE_VG_CHUNK) it must agree with
    the one-chunk result (frames are independent; the backward GEMM's summation order follows the tile's CTA, so to
    rounding), (b) tanh takes the fused forward + stored activations, other activations the layered recompute -- both
    must match the fp64 oracle."""
    monkeypatch.setenv("MOLANN_B200_WIDE", "1")
    spec = S.get_spec("C3s")
    spec.activation = activation
    model, _ = S.build_model(spec, init_seed=5)
    nl = len(spec.layer_dims) - 1
    ws, bs = _weights(model, nl)
    L = 777
    x = S.make_frames(spec, L, seed=21)
    cot = torch.randn(L, spec.out_dim(), generator=torch.Generator().manual_seed(8))
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
    model = model.cuda()
    y, gx = model.value_and_grad(dev(x), cot.cuda())
    assert_parity(y.cpu(), y64, y32, TOL, activation + " y")
    assert_parity(gx.cpu(), gx64, gx32, TOL, activation + " gx")
    monkeypatch.setenv("MOLANN_B200_WIDE_VG_CHUNK", "256")          # 777 frames = 3 chunks + a ragged one
    yc, gxc = model.value_and_grad(dev(x), cot.cuda())
    assert float((yc - y).abs().max()) <= 2e-6 * max(1.0, float(y.abs().max()))
    assert float((gxc - gx).abs().max()) <= 2e-5 * max(1e-6, float(gx.abs().max()))
    assert_parity(gxc.cpu(), gx64, gx32, TOL, activation + " gx (chunked)")
