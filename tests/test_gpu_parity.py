"""Parity of the CUDA path with the oracle / the reference's golden vectors (run on a B200: -m gpu).

Every case goes through the C ABI (ctypes, helpers.CPlan) and/or the drop-in module API (which calls the
same C ABI through the torch shim).  Acceptance rule: per frame,
|new - ref64| / |ref64| <= max(1e-5, 2 |ref32 - ref64| / |ref64|) in max-abs norm; integer side exact."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import (ROOT, S, CPlan, assert_parity, frame_rel_err, golden, golden_weights, oracle_model,
                     oracle_preprocess, oracle_value_and_grad, spec_program, R)

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


def test_native_library_is_loaded_and_counts_launches():
    from molann_b200 import _lib
    import molann_b200.ann  # noqa: F401
    maps = open("/proc/self/maps").read()
    assert "libmolann_b200.so" in maps and "libmolann_b200_torch.so" in maps
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    before = _lib.launch_count()
    model(S.make_frames(spec, 300, device="cuda"))
    torch.cuda.synchronize()
    assert _lib.launch_count() == before + 1          # ONE fused kernel for align + features + MLP
    assert int(torch.ops.molann_b200.launch_count()) == _lib.launch_count()


@pytest.mark.parametrize("path", ["auto", "tc_single", "ffma", "general"])
@pytest.mark.parametrize("name", ["C1", "C2", "C3s"])
def test_cabi_against_reference_goldens(name, path, monkeypatch):
    """auto = best fused kernel (warp-specialised tcgen05 3xTF32 pipeline where eligible), tc_single = the
    single-role tcgen05 kernel, ffma = fused CUDA-core kernel, general = warp-per-frame geometry + layered GEMMs."""
    if path == "general":
        monkeypatch.setenv("MOLANN_B200_PATH", "0")
    if path == "ffma":
        monkeypatch.setenv("MOLANN_B200_TC", "0")
    if path == "tc_single":
        monkeypatch.setenv("MOLANN_B200_WS", "0")
    spec = S.get_spec(name)
    g = golden("config_" + name)
    ws, bs = golden_weights(g, len(spec.layer_dims) - 1)
    plan = CPlan(spec, ws, bs)
    want = 0 if path == "general" else 1
    import ctypes
    assert plan.lib.molann_b200_path_for(ctypes.byref(plan.p), 0) == want
    if name == "C2":
        fam = plan.lib.molann_b200_kernel_family(ctypes.byref(plan.p), 0)
        assert fam == {"auto": 2, "tc_single": 2, "ffma": 1, "general": 0}[path]
    x = dev(g["x"])
    y = plan.forward(x)
    assert_parity(y.cpu(), g["y64"], g["y32"], TOL, "%s y" % name)
    gx = plan.backward(x, dev(g["cot"]))
    assert_parity(gx.cpu(), g["gx64"], g["gx32"], TOL, "%s gx" % name)
    feat = plan.preprocess_forward(x)
    assert_parity(feat.cpu(), g["feat64"], g["feat32"], TOL, "%s feat" % name)
    gxf = plan.preprocess_backward(x, dev(g["cotf"]))
    assert_parity(gxf.cpu(), g["gxf64"], g["gxf32"], TOL, "%s gxf" % name)


@pytest.mark.parametrize("name", ["C1", "C2", "C3s"])
def test_parameter_gradients_against_goldens(name):
    spec = S.get_spec(name)
    g = golden("config_" + name)
    nl = len(spec.layer_dims) - 1
    ws, bs = golden_weights(g, nl)
    plan = CPlan(spec, ws, bs)
    x = dev(g["x"])
    gx, gW, gb = plan.backward(x, dev(g["cot"]), want_params=True)
    assert_parity(gx.cpu(), g["gx64"], g["gx32"], TOL, "gx (general path)")
    for k in range(nl):
        rw = torch.from_numpy(g["gp64::ann_layers.%dth_layer.weight" % (k + 1)])
        rb = torch.from_numpy(g["gp64::ann_layers.%dth_layer.bias" % (k + 1)])
        assert float((gW[k].cpu().double() - rw).abs().max() / rw.abs().max()) < 2e-5
        assert float((gb[k].cpu().double() - rb).abs().max() / rb.abs().max()) < 2e-5


def test_fixture_known_answers_through_modules():
    """SURVEY App. C on the 22-atom fixture, through the drop-in classes."""
    from molann_b200.ann import AlignmentLayer, FeatureLayer
    from molann_b200.atomgroup import Universe
    from molann_b200.feature import Feature
    g = golden("fixture")
    u = Universe(g["x"])
    x = dev(g["x"]).unsqueeze(0)
    one = lambda *ids: u.select_ix([i - 1 for i in ids])
    hist = [("d1", "dihedral", (5, 7, 9, 15)), ("d2", "dihedral", (7, 9, 15, 17)), ("b1", "bond", (2, 5)),
            ("b2", "bond", (5, 6)), ("a1", "angle", (20, 19, 21)), ("a2", "angle", (16, 15, 17))]
    feats = [Feature(n, t, one(*ids)) for n, t, ids in hist]
    out = FeatureLayer(feats, u.atoms, use_angle_value=False).cuda()(x).cpu().numpy()
    np.testing.assert_allclose(out, g["hist_cs"], atol=2e-6)
    out = FeatureLayer(feats, u.atoms, use_angle_value=True).cuda()(x).cpu().numpy()
    np.testing.assert_allclose(out, g["hist_angle"], atol=2e-6)
    extra = [(1, 2, 5, 6), (3, 2, 5, 7), (10, 9, 11, 12), (9, 15, 17, 19)]
    fe = [Feature("e", "dihedral", one(*ids)) for ids in extra]
    out = FeatureLayer(fe, u.atoms).cuda()(x).cpu().numpy()
    np.testing.assert_allclose(out, g["extra_dihedrals"], atol=2e-6)
    inp = one(1, 2, 3, 4, 5)
    fl = FeatureLayer([Feature("n", "dihedral", one(1, 2, 3, 4)), Feature("n", "bond", one(1, 3)),
                       Feature("n", "angle", one(1, 2, 3))], inp).cuda()
    out = fl(dev(inp.positions).unsqueeze(0)).cpu().numpy()
    np.testing.assert_allclose(out, g["test_feature_layer"], atol=2e-6)
    al = AlignmentLayer(one(1, 2, 5), u.atoms).cuda()
    np.testing.assert_allclose(al(x).cpu().numpy(), g["align_self"], atol=3e-6)
    np.testing.assert_allclose(al(dev(g["align_moved_x"]).unsqueeze(0)).cpu().numpy(), g["align_moved"], atol=5e-6)


@pytest.mark.parametrize("name", ["C1", "C2", "C3s"])
def test_module_api_forward_and_autograd(name):
    spec = S.get_spec(name)
    g = golden("config_" + name)
    model, _ = S.build_model(spec)
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=True)
    model = model.cuda()
    x = dev(g["x"]).requires_grad_(True)
    y = model(x)
    assert_parity(y.detach().cpu(), g["y64"], g["y32"], TOL, "module y")
    (gx,) = torch.autograd.grad(y, x, dev(g["cot"]))
    assert_parity(gx.cpu(), g["gx64"], g["gx32"], TOL, "module gx")
    # preprocessing layer alone, and parameter gradients through autograd (training path)
    pp = model.get_preprocessing_layer()
    x2 = dev(g["x"]).requires_grad_(True)
    f = pp(x2)
    assert_parity(f.detach().cpu(), g["feat64"], g["feat32"], TOL, "module feat")
    (gxf,) = torch.autograd.grad(f, x2, dev(g["cotf"]))
    assert_parity(gxf.cpu(), g["gxf64"], g["gxf32"], TOL, "module gxf")
    model.zero_grad()
    (model(dev(g["x"])) * dev(g["cot"])).sum().backward()
    for k, p in model.named_parameters():
        ref = torch.from_numpy(g["gp64::" + k])
        assert float((p.grad.cpu().double() - ref).abs().max() / ref.abs().max()) < 2e-5, k


@pytest.mark.parametrize("name", ["C1", "C2", "C3s"])
def test_value_and_grad_single_pass(name):
    """The biasing-force entry point: (y, d<cot,y>/dx) from ONE launch where the plan is eligible."""
    from molann_b200 import _lib
    spec = S.get_spec(name)
    g = golden("config_" + name)
    model, _ = S.build_model(spec)
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=True)
    model = model.cuda()
    x, cot = dev(g["x"]), dev(g["cot"])
    before = _lib.launch_count()
    y, gx = model.value_and_grad(x, cot)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - before
    assert_parity(y.cpu(), g["y64"], g["y32"], TOL, "vg y")
    assert_parity(gx.cpu(), g["gx64"], g["gx32"], TOL, "vg gx")
    if name == "C2":
        assert launches == 1
    ys, gxs = torch.jit.script(model).value_and_grad(x, cot)
    assert torch.equal(ys, y) and torch.equal(gxs, gx)


@pytest.mark.parametrize("tile", ["1", "0"])
def test_alignment_layer_standalone_and_composition(tile, monkeypatch):
    """AlignmentLayer alone (fwd + autograd) and the unfused composition feature(align(x)); thread-per-frame tile
    kernels (1, the default for small frames) and warp-per-frame kernels (0)."""
    monkeypatch.setenv("MOLANN_B200_TILE", tile)
    spec = S.get_spec("C2")
    g = golden("config_C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    pp = model.get_preprocessing_layer()
    aidx, ref, feats, _, _ = spec_program(spec)
    x = dev(g["x"]).requires_grad_(True)
    z = pp.align_layer(x)
    x64 = torch.from_numpy(g["x"]).double().requires_grad_(True)
    z64 = R.align_forward(x64, aidx, ref.double())
    assert_parity(z.detach().cpu(), z64.detach(), None, TOL, "align z")
    cot = torch.randn(z.shape, generator=torch.Generator().manual_seed(3))
    (gz,) = torch.autograd.grad(z, x, cot.cuda())
    (gz64,) = torch.autograd.grad((z64 * cot.double()).sum(), x64)
    assert_parity(gz.cpu(), gz64, None, TOL, "align gx")
    f2 = pp.feature_layer(pp.align_layer(dev(g["x"])))
    assert_parity(f2.cpu(), g["feat64"], g["feat32"], TOL, "composed feat")
    # stand-alone preprocessing on ragged batches (partial last tile, 4-byte-offset base pointer)
    x_all = torch.from_numpy(g["x"])
    flat = torch.empty(x_all.numel() + 1, device="cuda")
    for L, shift in ((1, 0), (129, 1), (x_all.shape[0], 0)):
        xd = flat[shift:shift + L * x_all[0].numel()].view(L, *x_all.shape[1:])
        xd.copy_(x_all[:L])
        xd = xd.detach().requires_grad_(True)
        f = pp(xd)
        assert_parity(f.detach().cpu(), g["feat64"][:L], g["feat32"][:L], TOL, "feat L=%d" % L)
        (gxf,) = torch.autograd.grad(f, xd, dev(g["cotf"][:L]))
        assert_parity(gxf.cpu(), g["gxf64"][:L], g["gxf32"][:L], TOL, "gxf L=%d" % L)
        zz = pp.align_layer(xd)
        assert_parity(zz.detach().cpu(), z64.detach()[:L], None, TOL, "align z L=%d" % L)


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 63, 64, 65, 127, 128, 129, 255, 256, 257, 385, 1000])
def test_ragged_frame_counts(L):
    """Partial tiles, L < tile, L == 3 (the reference's torch.cross quirk case is defined by intent)."""
    for name in ("C1", "C2"):
        spec = S.get_spec(name)
        g = golden("config_" + name)
        ws, bs = golden_weights(g, len(spec.layer_dims) - 1)
        plan = CPlan(spec, ws, bs)
        x = S.make_frames(spec, L, seed=900 + L)
        cot = torch.randn(L, spec.out_dim(), generator=torch.Generator().manual_seed(L))
        y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
        y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
        assert_parity(plan.forward(dev(x)).cpu(), y64, y32, TOL, "%s L=%d y" % (name, L))
        assert_parity(plan.backward(dev(x), dev(cot)).cpu(), gx64, gx32, TOL, "%s L=%d gx" % (name, L))


def test_slices_of_a_batch_are_bitwise_identical():
    """A frame's output must not depend on where its batch was cut (offset views are 8 / 4 bytes off the
    16-byte grid, ragged tails, single frames): shards, chunks and re-batched frames reproduce bit for bit."""
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, 3000, device="cuda", seed=77)
    flat = torch.zeros(3000 * 66 + 3, device="cuda")
    with torch.no_grad():
        y = model(x)
        for a, n in ((1, 128), (3, 1000), (2, 129), (127, 1), (129, 257), (5, 2995)):
            assert torch.equal(model(x[a:a + n]), y[a:a + n]), (a, n)
        for shift in (1, 2, 3):                       # 4-, 8-, 12-byte misaligned copies of the whole batch
            xv = flat[shift:shift + 3000 * 66].view(3000, 22, 3)
            xv.copy_(x)
            assert xv.data_ptr() % 16 == 4 * shift
            assert torch.equal(model(xv), y), shift


def test_empty_batch():
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    y = model(torch.zeros(0, 22, 3, device="cuda"))
    assert tuple(y.shape) == (0, 2)


def test_unaligned_input_pointer_takes_the_non_tma_path():
    spec = S.get_spec("C2")
    g = golden("config_C2")
    ws, bs = golden_weights(g, 3)
    plan = CPlan(spec, ws, bs)
    base = torch.zeros(1 + g["x"].size, device="cuda")
    xv = base[1:].view(g["x"].shape)                  # 4-byte aligned only
    xv.copy_(dev(g["x"]))
    assert xv.data_ptr() % 16 != 0
    # the forward kernel bulk-copies from the 16-byte boundary below each tile: same kernel, same bits
    assert torch.equal(plan.forward(xv), plan.forward(dev(g["x"])))
    assert_parity(plan.forward(xv).cpu(), g["y64"], g["y32"], TOL, "unaligned y")
    assert_parity(plan.backward(xv, dev(g["cot"])).cpu(), g["gx64"], g["gx32"], TOL, "unaligned gx")


@pytest.mark.parametrize("act", ["relu", "sigmoid"])
def test_other_activations(act):
    spec = S.get_spec("C2")
    spec.activation = act
    model, _ = S.build_model(spec, init_seed=5)
    sd = model.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k] for k in (1, 2, 3)]
    bs = [sd["ann_layers.%dth_layer.bias" % k] for k in (1, 2, 3)]
    x = S.make_frames(spec, 333, seed=8)
    cot = torch.randn(333, 2, generator=torch.Generator().manual_seed(1))
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
    model = model.cuda()
    xd = dev(x).requires_grad_(True)
    y = model(xd)
    (gx,) = torch.autograd.grad(y, xd, cot.cuda())
    assert_parity(y.detach().cpu(), y64, y32, TOL, act + " y")
    if act == "relu":       # kinks: frames with a pre-activation within fp32 noise of 0 may flip a unit
        err = frame_rel_err(gx.cpu(), gx64)
        assert float((err > TOL).float().mean()) < 0.02
    else:
        assert_parity(gx.cpu(), gx64, gx32, TOL, act + " gx")


@pytest.mark.parametrize("ws", ["1", "0"])
@pytest.mark.parametrize("variant", ["one_hidden", "three_outputs", "narrow_wide", "mixed_program", "no_alignment",
                                     "sigmoid_mixed"])
def test_tensor_core_kernel_variants(variant, ws, monkeypatch):
    """Shapes of the small-system class other than C2: one hidden layer, hidden width that is not a multiple
    of 64, up to 8 outputs, feature programs that mix position entries with internal coordinates (unrolled
    position path + interpreter), no alignment layer.  ws=1: warp-specialised pipeline, ws=0: single-role kernel."""
    monkeypatch.setenv("MOLANN_B200_WS", ws)
    spec = S.get_spec("C2")
    if variant == "one_hidden":
        spec.layer_dims = [30, 48, 2]
    elif variant == "three_outputs":
        spec.layer_dims = [30, 64, 64, 3]
    elif variant == "narrow_wide":
        spec.layer_dims = [30, 16, 64, 8]
    elif variant in ("mixed_program", "sigmoid_mixed"):
        spec.features = [("p", "position", [1, 4, 5, 6, 8]), ("d", "dihedral", [4, 6, 8, 14]), ("b", "bond", [8, 10]),
                         ("a", "angle", [6, 8, 14]), ("p2", "position", [16, 18])]
        spec.layer_dims = [spec.feature_dim(), 32, 40, 2]
        if variant == "sigmoid_mixed":
            spec.activation = "sigmoid"
    elif variant == "no_alignment":
        spec.align_ix = None
        spec.trans_sigma, spec.rotate = 0.0, False
    model, _ = S.build_model(spec, init_seed=3)
    sd = model.state_dict()
    nl = len(spec.layer_dims) - 1
    ws_ = [sd["ann_layers.%dth_layer.weight" % (k + 1)] for k in range(nl)]
    bs_ = [sd["ann_layers.%dth_layer.bias" % (k + 1)] for k in range(nl)]
    L = 777
    x = S.make_frames(spec, L, seed=12)
    cot = torch.randn(L, spec.out_dim(), generator=torch.Generator().manual_seed(6))
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws_, bs_), x, cot)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws_, bs_, torch.float32), x, cot, torch.float32)
    model = model.cuda()
    from molann_b200 import _lib
    before = _lib.launch_count()
    with torch.no_grad():
        y = model(dev(x))
    assert _lib.launch_count() == before + 1
    assert_parity(y.cpu(), y64, y32, TOL, variant + " y")
    y2, gx = model.value_and_grad(dev(x), dev(cot))
    assert_parity(y2.cpu(), y64, y32, TOL, variant + " y (value_and_grad)")
    assert_parity(gx.cpu(), gx64, gx32, TOL, variant + " gx")


def test_mixed_feature_program_with_alignment():
    """All four feature types + alignment in one layer, angle values on/off, shuffled input group."""
    from molann_b200.ann import AlignmentLayer, FeatureLayer, PreprocessingANN
    from molann_b200.atomgroup import Universe
    from molann_b200.feature import Feature
    rng = np.random.RandomState(21)
    u = Universe(S.ala2_positions())
    inp_ix = rng.permutation(22)[:17]
    inp = u.select_ix(inp_ix)
    pick = lambda m: u.select_ix(rng.permutation(inp_ix)[:m])
    feats = [Feature("d", "dihedral", pick(4)), Feature("p", "position", pick(5)), Feature("b", "bond", pick(2)),
             Feature("a", "angle", pick(3)), Feature("d2", "dihedral", pick(4)), Feature("p2", "position", pick(2))]
    al_group = pick(6)
    for ua in (False, True):
        pp = PreprocessingANN(AlignmentLayer(al_group, inp), FeatureLayer(feats, inp, use_angle_value=ua)).cuda()
        base = torch.from_numpy(u.atoms.positions[inp_ix])
        x = base.unsqueeze(0) + 0.25 * torch.randn(500, 17, 3, generator=torch.Generator().manual_seed(2))
        x = torch.bmm(x, S.random_rotations(500, torch.Generator().manual_seed(3), "cpu")) + 7.0
        inp_list = inp_ix.tolist()
        fl = [(f.get_type_id(), [inp_list.index(i) for i in f.atom_group.ix]) for f in feats]
        aidx = [inp_list.index(i) for i in al_group.ix]
        ref = torch.from_numpy(al_group.positions)
        ref = ref - ref.mean(0)
        fn = lambda xx: R.preprocess_forward(xx, aidx, ref.double(), fl, ua)
        cot = torch.randn(500, pp.output_dimension(), generator=torch.Generator().manual_seed(4))
        f64, gx64 = oracle_value_and_grad(fn, x, cot)
        xd = dev(x).requires_grad_(True)
        f = pp(xd)
        (gx,) = torch.autograd.grad(f, xd, cot.cuda())
        assert_parity(f.detach().cpu(), f64, None, TOL, "mixed feat ua=%s" % ua)
        err = frame_rel_err(gx.cpu(), gx64)          # acos / atan2 derivative is ill-conditioned near 0, pi
        assert float(err.median()) < 1e-5 and float((err > 1e-4).float().mean()) < 0.02


def test_large_coordinate_offset():
    spec = S.get_spec("C2")
    g = golden("config_C2")
    ws, bs = golden_weights(g, 3)
    plan = CPlan(spec, ws, bs)
    x = torch.from_numpy(g["x"]) + 1000.0
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, torch.from_numpy(g["cot"]))
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, torch.from_numpy(g["cot"]),
                                      torch.float32)
    assert_parity(plan.forward(dev(x)).cpu(), y64, y32, 1e-4, "offset y")
    assert_parity(plan.backward(dev(x), dev(g["cot"])).cpu(), gx64, gx32, 1e-4, "offset gx")


def test_full_size_properties_c2():
    """BASELINE size (1 Mi frames): size-independent properties + a sampled oracle check."""
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    L = 1 << 20
    x = S.make_frames(spec, L, device="cuda")
    with torch.no_grad():
        y = model(x)
        assert torch.isfinite(y).all()
        # frames are independent: any sub-batch / permutation gives bit-identical rows
        perm = torch.randperm(L, device="cuda")
        assert torch.equal(model(x[perm].contiguous()), y[perm])
        assert torch.equal(model(x[12345:12345 + 777].contiguous()), y[12345:12345 + 777])
        # rigid-motion invariance of the aligned model
        Rm = S.random_rotations(L, torch.Generator(device="cuda").manual_seed(7), "cuda")
        x2 = torch.bmm(x, Rm) + 3.0
        assert float((model(x2) - y).abs().max()) < 2e-4
    cot = torch.randn(L, 2, device="cuda")
    xg = x.clone().requires_grad_(True)
    (g1,) = torch.autograd.grad(model(xg), xg, cot)
    (g2,) = torch.autograd.grad(model(xg), xg, 2.0 * cot)
    assert torch.equal(g2, 2.0 * g1)                                   # linear in the cotangent, exactly
    assert float(g1.sum(dim=1).abs().max()) < 1e-3 * float(g1.abs().max())      # translation invariance
    sd = model.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k].cpu() for k in (1, 2, 3)]
    bs = [sd["ann_layers.%dth_layer.bias" % k].cpu() for k in (1, 2, 3)]
    idx = torch.randint(0, L, (2000,), generator=torch.Generator().manual_seed(0))
    xs, cs = x[idx.cuda()].cpu(), cot[idx.cuda()].cpu()
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), xs, cs)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), xs, cs, torch.float32)
    assert_parity(y[idx.cuda()].cpu(), y64, y32, TOL, "full-size y sample")
    assert_parity(g1[idx.cuda()].cpu(), gx64, gx32, TOL, "full-size gx sample")


@pytest.mark.parametrize("name", ["C3", "C5"])
def test_bench_size_properties_big_systems(name):
    """BASELINE configs[2] / [4] at the bench's batch size (32768 x 2000 atoms / 16384 x 5000 atoms), through
    size-independent properties: frames independent (slices equal to rounding), invariance of the aligned model under a
    rigid motion of the input, the coordinate gradient rotating WITH the frame, linear in the cotangent, summing to
    zero over atoms (translation invariance), zero on unreferenced atoms; plus a 48-frame sample against the fp64
    oracle."""
    spec = S.get_spec(name)
    L = spec.default_frames
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, L, device="cuda", seed=9)
    cot = torch.randn(L, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    y, g1 = model.value_and_grad(x, cot)
    assert torch.isfinite(y).all() and torch.isfinite(g1).all()
    # frames are independent; on the layered path a row's K-summation order follows the CTA that owns its tile (the
    # GEMM rotates the chunk order per CTA), so a slice agrees to rounding, not bitwise like the fused kernels
    ys, gs = model.value_and_grad(x[1000:1000 + 333].contiguous(), cot[1000:1000 + 333].contiguous())
    assert float((ys - y[1000:1333]).abs().max()) <= 2e-6 * max(1.0, float(y.abs().max()))
    assert float((gs - g1[1000:1333]).abs().max()) <= 1e-5 * float(g1.abs().max())
    Ls = 2048
    Rm = S.random_rotations(Ls, torch.Generator(device="cuda").manual_seed(7), "cuda")
    x2 = (torch.bmm(x[:Ls], Rm) + 5.0).contiguous()
    y2, g2 = model.value_and_grad(x2, cot[:Ls].contiguous())
    scale_y = float(y[:Ls].abs().max())
    assert float((y2 - y[:Ls]).abs().max()) < 3e-4 * max(1.0, scale_y)
    g1r = torch.bmm(g1[:Ls], Rm)                                   # d/dx of an invariant function rotates with x
    cov = (g2 - g1r).abs().amax(dim=(1, 2)) / float(g1[:Ls].abs().max())
    assert float(cov.median()) < 5e-5 and float(cov.max()) < 3e-3, (float(cov.median()), float(cov.max()))
    _, g3 = model.value_and_grad(x[:Ls].contiguous(), (2.0 * cot[:Ls]).contiguous())
    assert float((g3 - 2.0 * g1[:Ls]).abs().max()) <= 1e-6 * float(g1[:Ls].abs().max())
    assert float(g1.sum(dim=1).abs().max()) < 2e-3 * float(g1.abs().max())
    used = set(spec.align_ix)
    for _, _, ix in spec.features:
        used.update(ix)
    unused = torch.tensor(sorted(set(range(spec.n_inp)) - used), device="cuda")
    assert float(g1[:, unused].abs().max()) == 0.0
    sd = model.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k].cpu() for k in (1, 2, 3)]
    bs = [sd["ann_layers.%dth_layer.bias" % k].cpu() for k in (1, 2, 3)]
    idx = torch.randint(0, L, (48,), generator=torch.Generator().manual_seed(0))
    xs, cs = x[idx.cuda()].cpu(), cot[idx.cuda()].cpu()
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), xs, cs)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), xs, cs, torch.float32)
    assert_parity(y[idx.cuda()].cpu(), y64, y32, TOL, name + " bench-size y sample")
    assert_parity(g1[idx.cuda()].cpu(), gx64, gx32, TOL, name + " bench-size gx sample")


@pytest.mark.parametrize("kernels", ["tensor_core_staged", "tensor_core_warp_staged", "ffma_gather"])
def test_c3_full_width_general_path(kernels, monkeypatch):
    """n = 2000 atoms, d = 800 -> [800,256,128,2]: warp-per-frame geometry + layered GEMMs.
    tensor_core_staged = smem-staged preprocess kernels + tcgen05 3xTF32 GEMMs with segmented accumulation
    (K = 800 is where the tensor core's accumulator rounding shows); ffma_gather = the CUDA-core kernels."""
    if kernels == "ffma_gather":
        monkeypatch.setenv("MOLANN_B200_GEMM_TC", "0")
        monkeypatch.setenv("MOLANN_B200_STAGED", "0")
    elif kernels == "tensor_core_warp_staged":
        monkeypatch.setenv("MOLANN_B200_STAGED", "1")
    spec = S.get_spec("C3")
    model, _ = S.build_model(spec)
    sd = model.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k] for k in (1, 2, 3)]
    bs = [sd["ann_layers.%dth_layer.bias" % k] for k in (1, 2, 3)]
    L = 203                                      # two row tiles, the second ragged; odd frame count
    x = S.make_frames(spec, L, seed=17)
    cot = torch.randn(L, 2, generator=torch.Generator().manual_seed(5))
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
    model = model.cuda()
    xd = dev(x).requires_grad_(True)
    y = model(xd)
    (gx,) = torch.autograd.grad(y, xd, cot.cuda())
    assert_parity(y.detach().cpu(), y64, y32, TOL, "C3 y")
    assert_parity(gx.cpu(), gx64, gx32, TOL, "C3 gx")
    assert int((gx.cpu() != 0).any(dim=2).sum(dim=1).max()) <= 200 + 400       # dense row, sparse support


@pytest.mark.parametrize("staged", ["3", "2", "1", "0"])
@pytest.mark.parametrize("with_align", [True, False])
def test_big_frame_preprocess_kernel_sets(staged, with_align, monkeypatch):
    """Preprocessing of a 303-atom system (rows of 3636 bytes: every other row starts off the 16-byte grid, the last
    one takes the non-bulk path) with all four feature types, several contributions per atom, 77 frames: the
    block-per-frame role pipeline in both directions (3) or for the backward only (2, the default), the
    warp-per-frame staged kernels (1) and the gather kernels (0)."""
    from molann_b200.ann import AlignmentLayer, FeatureLayer, PreprocessingANN
    from molann_b200.atomgroup import Universe
    from molann_b200.feature import Feature
    monkeypatch.setenv("MOLANN_B200_STAGED", staged)
    n, L = 303, 77
    rng = np.random.RandomState(5)
    u = Universe(S.chain_positions(n, 11))
    inp = u.select_ix(np.arange(n))
    sel = np.arange(0, n, 7)
    feats = [Feature("p", "position", u.select_ix(sel))]
    for k in range(40):
        a = 7 * k
        feats.append(Feature("d%d" % k, "dihedral", u.select_ix([a, a + 1, a + 2, a + 3])))
        feats.append(Feature("b%d" % k, "bond", u.select_ix([a + 3, a + 4])))
        feats.append(Feature("a%d" % k, "angle", u.select_ix([a + 2, a + 4, a + 5])))
    feats.append(Feature("p2", "position", u.select_ix(rng.permutation(n)[:30])))
    al = AlignmentLayer(u.select_ix(sel), inp) if with_align else None
    pp = PreprocessingANN(al, FeatureLayer(feats, inp, use_angle_value=False)).cuda()
    base = torch.from_numpy(u.atoms.positions)
    g = torch.Generator().manual_seed(8)
    x = base.unsqueeze(0) + 0.2 * torch.randn(L, n, 3, generator=g)
    x = torch.bmm(x, S.random_rotations(L, g, "cpu")) + 11.0
    fl = [(f.get_type_id(), list(f.atom_group.ix)) for f in feats]
    ref = torch.from_numpy(u.atoms.positions[sel])
    ref = (ref - ref.mean(0)).double()
    fn = lambda xx: R.preprocess_forward(xx, list(sel) if with_align else None, ref if with_align else None, fl, False)
    cot = torch.randn(L, pp.output_dimension(), generator=g)
    f64, gx64 = oracle_value_and_grad(fn, x, cot)
    ref32 = ref.float()
    fn32 = lambda xx: R.preprocess_forward(xx, list(sel) if with_align else None, ref32 if with_align else None, fl,
                                           False)
    f32, gx32 = oracle_value_and_grad(fn32, x, cot, torch.float32)
    flat = torch.empty(L * n * 3 + 1, device="cuda")
    for shift in (0, 1):                                   # 16-byte aligned base, then a base 4 bytes off
        xd = flat[shift:shift + L * n * 3].view(L, n, 3)
        xd.copy_(x)
        xd = xd.detach().requires_grad_(True)
        f = pp(xd)
        (gx,) = torch.autograd.grad(f, xd, cot.cuda())
        assert_parity(f.detach().cpu(), f64, f32, TOL, "big-frame features staged=%s shift=%d" % (staged, shift))
        assert_parity(gx.cpu(), gx64, gx32, TOL, "big-frame gx staged=%s shift=%d" % (staged, shift))
        for Ls in (1, 2):                                  # fewer frames than ring stages / CTAs
            xs = xd.detach()[:Ls].contiguous().requires_grad_(True)
            fs = pp(xs)
            (gs,) = torch.autograd.grad(fs, xs, cot[:Ls].cuda())
            assert_parity(fs.detach().cpu(), f64[:Ls], f32[:Ls], TOL, "big-frame features L=%d" % Ls)
            assert_parity(gs.cpu(), gx64[:Ls], gx32[:Ls], TOL, "big-frame gx L=%d" % Ls)


def test_c5_shape_parity():
    """BASELINE configs[4] shape (5000 atoms, 500-atom selection, d = 2000 -> [2000,256,128,2]) on a few frames."""
    spec = S.get_spec("C5")
    model, _ = S.build_model(spec)
    sd = model.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k] for k in (1, 2, 3)]
    bs = [sd["ann_layers.%dth_layer.bias" % k] for k in (1, 2, 3)]
    L = 131
    x = S.make_frames(spec, L, seed=23)
    cot = torch.randn(L, 2, generator=torch.Generator().manual_seed(6))
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
    model = model.cuda()
    y, gx = model.value_and_grad(dev(x), cot.cuda())
    assert_parity(y.cpu(), y64, y32, TOL, "C5 y")
    assert_parity(gx.cpu(), gx64, gx32, TOL, "C5 gx")


def test_c4_training_step_gradients_and_sgd():
    """BASELINE configs[3]: autoencoder step -- encoder = C2 MolANN through the fused kernels, decoder
    = create_sequential_nn([2,64,64,30]), MSE against the preprocessing output.  Step-1 parameter gradients and the
    loss against the fp64 oracle on a 4096-frame batch, then two SGD steps must lower the loss."""
    import copy
    sys.path.insert(0, ROOT)
    from bench import c4_models
    from molann_b200.train import AutoencoderStep
    spec, enc, dec = c4_models()
    L = 4096
    x = S.make_frames(spec, L, seed=404)
    sd = enc.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k].double().clone().requires_grad_(True) for k in (1, 2, 3)]
    bs = [sd["ann_layers.%dth_layer.bias" % k].double().clone().requires_grad_(True) for k in (1, 2, 3)]
    dec64 = copy.deepcopy(dec).double()
    target = oracle_preprocess(spec)(x)
    recon = dec64(oracle_model(spec, ws, bs)(x))
    loss64 = ((recon - target) ** 2).sum() / (L * target.shape[1])
    loss64.backward()
    ref = {}
    for k in (1, 2, 3):
        ref["enc.%d.w" % k], ref["enc.%d.b" % k] = ws[k - 1].grad, bs[k - 1].grad
    for name, p in dec64.named_parameters():
        ref["dec." + name] = p.grad
    enc, dec = enc.cuda(), dec.cuda()
    trainer = AutoencoderStep(enc, dec, lr=1e-3, global_frames=L)
    loss = trainer.loss_and_grads(dev(x))
    assert abs(float(loss) - float(loss64)) < 1e-5 * abs(float(loss64))
    got = {}
    for k in (1, 2, 3):
        layer = getattr(enc.ann_layers, "%dth_layer" % k)
        got["enc.%d.w" % k], got["enc.%d.b" % k] = layer.weight.grad, layer.bias.grad
    for name, p in dec.named_parameters():
        got["dec." + name] = p.grad
    for name, r in ref.items():
        err = float((got[name].cpu().double() - r).abs().max() / r.abs().max())
        assert err < 2e-5, (name, err)
    l0 = float(trainer.step(dev(x)))
    trainer.lr = 0.05
    for _ in range(3):
        l1 = float(trainer.step(dev(x)))
    assert l1 < l0


def test_torchscript_roundtrip_in_fresh_process(tmp_path):
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, 777, device="cuda")
    cot = torch.randn(777, 2, device="cuda")
    xg = x.clone().requires_grad_(True)
    y = model(xg)
    (gx,) = torch.autograd.grad(y, xg, cot)
    scripted = torch.jit.script(model)
    assert torch.equal(scripted(x), y.detach())
    path = os.path.join(tmp_path, "molann.pt")
    scripted.save(path)
    torch.save({"x": x.cpu(), "cot": cot.cpu(), "y": y.detach().cpu(), "gx": gx.cpu()}, os.path.join(tmp_path, "io.pt"))
    code = """
import sys, torch
torch.ops.load_library(sys.argv[1])            # a consumer only needs the shim library, not the Python package
m = torch.jit.load(sys.argv[2]).cuda()
io = torch.load(sys.argv[3])
x = io['x'].cuda().requires_grad_(True)
y = m(x)
(gx,) = torch.autograd.grad(y, x, io['cot'].cuda())
assert torch.equal(y.detach().cpu(), io['y']), 'forward differs'
assert torch.equal(gx.cpu(), io['gx']), 'gradient differs'
print('ROUNDTRIP_OK')
"""
    lib = os.path.join(ROOT, "molann_b200", "libmolann_b200_torch.so")
    out = subprocess.run([sys.executable, "-c", code, lib, path, os.path.join(tmp_path, "io.pt")],
                         capture_output=True, text=True, timeout=600)
    assert "ROUNDTRIP_OK" in out.stdout, out.stdout + out.stderr


def test_host_pipeline_end_to_end():
    from molann_b200.stream import HostPipeline
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    L = 50000
    xh = S.make_frames(spec, L).pin_memory()
    yh = torch.empty(L, 2).pin_memory()
    pipe = HostPipeline(model, 22, 2, chunk_frames=8192)
    pipe.run(xh, yh)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = model(xh.cuda()).cpu()
    assert torch.equal(yh, ref)
    assert pipe.h2d_bytes == L * 22 * 12 and pipe.d2h_bytes == L * 2 * 4
    coth = torch.randn(L, 2).pin_memory()
    gxh = torch.empty(L, 22, 3).pin_memory()
    pipe.run(xh, yh, coth, gxh)
    torch.cuda.synchronize()
    xg = xh.cuda().requires_grad_(True)
    (gref,) = torch.autograd.grad(model(xg), xg, coth.cuda())
    assert torch.equal(gxh, gref.cpu())


def test_wrong_dtype_device_layout_raise():
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, 16, device="cuda")
    with pytest.raises(RuntimeError, match="float32"):
        model(x.double())
    with pytest.raises(RuntimeError, match="contiguous"):
        model(x.transpose(0, 1).contiguous().transpose(0, 1))
    cpu_model, _ = S.build_model(spec)
    with pytest.raises(RuntimeError, match="device"):
        cpu_model(x)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_inference_is_bitwise_identical():
    from molann_b200.shard import frame_range
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    x = S.make_frames(spec, 100001)
    with torch.no_grad():
        full = model.cuda(0)(x.cuda(0)).cpu()
        parts = []
        G = torch.cuda.device_count()
        for r in range(G):
            s, e = frame_range(x.shape[0], r, G)
            m = S.build_model(spec)[0].cuda(r)
            with torch.cuda.device(r):
                parts.append(m(x[s:e].cuda(r)).cpu())
    assert torch.equal(torch.cat(parts), full)
