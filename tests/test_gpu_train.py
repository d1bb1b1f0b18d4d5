"""Fused autoencoder training kernel (csrc/fused_train.cuh; include/molann_b200.h molann_b200_train_*) against the fp64
oracle: loss = mean((decoder(MolANN(x)) - PreprocessingANN(x))^2) built from the reference's own modules
(/root/reference/molann/ann.py:37-67 create_sequential_nn, :553-565 PreprocessingANN, :567-624 MolANN) and
differentiated by autograd w.r.t. every Linear parameter.  Tolerance: 2e-5 of the largest entry of each gradient tensor
(the sums run over up to 10^5 fp32 terms), 1e-5 relative on the loss."""
import copy
import dataclasses

import pytest
import torch

from helpers import S, oracle_model, oracle_preprocess

pytestmark = pytest.mark.gpu

ACT_IDS = {"tanh": 0, "relu": 1, "sigmoid": 2}


def make_pair(spec, dec_hidden, dec_act="tanh", seed=11):
    enc, _ = S.build_model(spec)
    torch.manual_seed(seed)
    create = S.default_api().create_sequential_nn
    dec = create([spec.out_dim()] + list(dec_hidden) + [spec.feature_dim()], S._ACTS[dec_act]())
    return enc, dec


def oracle_loss_and_grads(spec, enc, dec, x, n_global=None):
    """fp64 loss and flat gradient in torch parameter order (encoder W1, b1, ..., decoder W1, b1, ...)."""
    nl = len(spec.layer_dims) - 1
    sd = enc.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % (k + 1)].double().cpu().clone().requires_grad_(True) for k in range(nl)]
    bs = [sd["ann_layers.%dth_layer.bias" % (k + 1)].double().cpu().clone().requires_grad_(True) for k in range(nl)]
    dec64 = copy.deepcopy(dec).cpu().double()
    target = oracle_preprocess(spec)(x)
    recon = dec64(oracle_model(spec, ws, bs)(x))
    n = x.shape[0] if n_global is None else n_global
    loss = ((recon - target) ** 2).sum() / (n * target.shape[1])
    loss.backward()
    grads = []
    for w, b in zip(ws, bs):
        grads += [w.grad, b.grad]
    grads += [p.grad for p in dec64.parameters()]
    return float(loss.detach()), grads


def fused_flat(spec, enc, dec, x, dec_act, n_global=None):
    from molann_b200 import ann as A
    enc, dec = enc.cuda(), dec.cuda()
    pp, fl = enc.preprocessing_layer, enc.preprocessing_layer.feature_layer
    geo = (pp.align_layer._align_idx, pp.align_layer.ref_x) if enc._fused_align else (fl._no_idx, fl._no_ref)
    enc_params = [t.detach() for layer in enc.ann_layers if hasattr(layer, "weight") for t in (layer.weight, layer.bias)]
    dec_params = [t.detach() for layer in dec if hasattr(layer, "weight") for t in (layer.weight, layer.bias)]
    args = geo + (fl._entries, fl._dim, fl.use_angle_value, enc_params, enc._act_id, dec_params, ACT_IDS[dec_act])
    xd = x.cuda()
    assert torch.ops.molann_b200.train_eligible(xd, *args)
    n = x.shape[0] if n_global is None else n_global
    flat = torch.ops.molann_b200.train_loss_and_grads(xd, *args, 1.0 / (n * fl._dim))
    return flat, enc_params + dec_params


def check(flat, params, loss64, grads64, what):
    off = 0
    for k, (p, g) in enumerate(zip(params, grads64)):
        got = flat[off:off + p.numel()].view_as(p).cpu().double()
        off += p.numel()
        scale = float(g.abs().max())
        err = float((got - g).abs().max()) / max(scale, 1e-30)
        assert err < 2e-5, "%s: gradient tensor %d (shape %s) off by %.3e of its largest entry" % (
            what, k, tuple(p.shape), err)
    assert off + 1 == flat.numel()
    assert abs(float(flat[off]) - loss64) <= 1e-5 * abs(loss64), (what, float(flat[off]), loss64)


@pytest.mark.parametrize("L", [1, 127, 128, 129, 1000, 148 * 128 + 77])
def test_c4_loss_and_every_gradient_ragged_batches(L):
    """C4 pair (encoder = C2 MolANN, decoder [2,64,64,30]); batches that end inside a tile, fill exactly one, and give
    some CTAs one tile more than others."""
    spec = S.get_spec("C2")
    enc, dec = make_pair(spec, [64, 64])
    x = S.make_frames(spec, L, seed=1000 + L)
    loss64, grads64 = oracle_loss_and_grads(spec, enc, dec, x)
    flat, params = fused_flat(spec, enc, dec, x, "tanh")
    check(flat, params, loss64, grads64, "C4 L=%d" % L)


@pytest.mark.parametrize("enc_act,dec_act", [("sigmoid", "relu"), ("relu", "tanh"), ("tanh", "sigmoid")])
def test_activation_pairs(enc_act, dec_act):
    spec = dataclasses.replace(S.get_spec("C2"), activation=enc_act)
    enc, dec = make_pair(spec, [48, 20], dec_act)            # 20: the half-width work shape; 48: the 4 x 4 one
    x = S.make_frames(spec, 777, seed=5)
    loss64, grads64 = oracle_loss_and_grads(spec, enc, dec, x)
    flat, params = fused_flat(spec, enc, dec, x, dec_act)
    check(flat, params, loss64, grads64, "%s/%s" % (enc_act, dec_act))


def test_two_feature_system_without_alignment():
    """C1 (bond + dihedral angle value, no AlignmentLayer, d_feat = 2): every layer takes the narrow / thin shapes
    somewhere, the reconstruction target is two columns wide."""
    spec = S.get_spec("C1")
    enc, dec = make_pair(spec, [7])
    x = S.make_frames(spec, 1500, seed=9)
    loss64, grads64 = oracle_loss_and_grads(spec, enc, dec, x)
    flat, params = fused_flat(spec, enc, dec, x, "tanh")
    check(flat, params, loss64, grads64, "C1")


def test_global_batch_scale_and_determinism():
    """loss_scale = 1 / (global frames * d_feat): a shard's flat vector is its share of the global mean; two calls on
    the same inputs give the same bits (no atomics anywhere)."""
    spec = S.get_spec("C2")
    enc, dec = make_pair(spec, [64, 64])
    x = S.make_frames(spec, 3000, seed=77)
    loss64, grads64 = oracle_loss_and_grads(spec, enc, dec, x, n_global=12000)
    flat, params = fused_flat(spec, enc, dec, x, "tanh", n_global=12000)
    check(flat, params, loss64, grads64, "shard of a global batch")
    again, _ = fused_flat(spec, enc, dec, x, "tanh", n_global=12000)
    assert torch.equal(flat, again)


def test_empty_batch_and_sgd_apply():
    spec = S.get_spec("C2")
    enc, dec = make_pair(spec, [64, 64])
    flat, params = fused_flat(spec, enc, dec, S.make_frames(spec, 0, seed=1), "tanh", n_global=10)
    assert flat.numel() == 12577 and not bool(flat.any())
    flat = torch.randn(12577, device="cuda")
    before = [p.clone() for p in params]
    torch.ops.molann_b200.sgd_apply_(params, flat, 0.25)
    off = 0
    for p, q in zip(params, before):
        want = torch.addcmul(q, flat[off:off + q.numel()].view_as(q), torch.tensor(-0.25, device="cuda"))
        assert torch.allclose(p, want, rtol=0, atol=1e-7)
        off += q.numel()


def test_autoencoder_step_fused_and_composed_agree(monkeypatch):
    """AutoencoderStep picks the fused kernel for the C4 pair and the composed path (fused encoder kernels + library
    decoder) when told not to; both produce the same loss and gradients to fp32 accuracy.  A pair too large for one
    SM (reduced C3: 200 atoms, 80 features) falls back by itself."""
    from molann_b200.train import AutoencoderStep
    spec = S.get_spec("C2")
    x = S.make_frames(spec, 2048, seed=3).cuda()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MOLANN_B200_TRAIN_FUSED", mode)
        enc, dec = make_pair(spec, [64, 64])
        tr = AutoencoderStep(enc.cuda(), dec.cuda(), lr=1e-3, global_frames=2048)
        loss = float(tr.loss_and_grads(x))
        assert (tr._fused_args is not None) == (mode == "1")
        out[mode] = (loss, [p.grad.clone() for p in tr.params])
    assert abs(out["1"][0] - out["0"][0]) < 1e-5 * abs(out["0"][0])
    for a, b in zip(out["1"][1], out["0"][1]):
        assert float((a - b).abs().max()) < 4e-5 * float(b.abs().max())
    monkeypatch.setenv("MOLANN_B200_TRAIN_FUSED", "1")
    small = S.get_spec("C3s")
    enc, dec = make_pair(small, [24, 48])
    tr = AutoencoderStep(enc.cuda(), dec.cuda(), lr=1e-3)
    l0 = float(tr.step(S.make_frames(small, 256, seed=4).cuda()))
    assert tr._fused_args is None and l0 > 0


def test_two_gpu_peer_memory_allreduce_and_sgd():
    """The collective of the fused step (one-shot allreduce over NVLink peer memory + SGD in one kernel,
    molann_b200_allreduce_sgd) on two ranks: equals the full-batch gradient, trains like the NCCL route, keeps the
    replicas bit-identical, replays as a CUDA graph (tests/cuda/train_p2p_check.py)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(here, "cuda", "train_p2p_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "P2P_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_step_from_host_equals_step_on_device():
    """AutoencoderStep.step_from_host streams the pinned shard in pieces under the kernel; the pieces' flat vectors add
    up to the one-launch result (to fp32 summation order), ragged last piece included."""
    from molann_b200.train import AutoencoderStep
    spec = S.get_spec("C2")
    L = 5 * 1024 + 333
    xh = S.make_frames(spec, L, seed=21).pin_memory()
    out = {}
    for mode in ("device", "host"):
        enc, dec = make_pair(spec, [64, 64])
        tr = AutoencoderStep(enc.cuda(), dec.cuda(), lr=0.05, global_frames=L)
        losses = []
        for _ in range(3):
            losses.append(float(tr.step(xh.cuda()) if mode == "device" else tr.step_from_host(xh, chunks=3)))
        torch.cuda.synchronize()
        out[mode] = (losses, [p.detach().clone() for p in tr.params])
    for a, b in zip(out["host"][0], out["device"][0]):
        assert abs(a - b) < 1e-5 * abs(b), out
    assert out["device"][0][-1] < out["device"][0][0]
    for p, q in zip(out["host"][1], out["device"][1]):
        assert float((p - q).abs().max()) < 1e-5 * float(q.abs().max())
