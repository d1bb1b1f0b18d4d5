"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding and the single flat
gradient allreduce of the training configuration.  The oracle stands in for the kernels as compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import S, golden, golden_weights, oracle_model
from molann_b200.shard import allreduce_flat_grads, frame_range, shard_sizes


def test_frame_range_partitions_exactly():
    for L in (0, 1, 7, 8, 1000, 1 << 20, 10 ** 8 + 3):
        for G in (1, 2, 3, 4, 8):
            ranges = [frame_range(L, r, G) for r in range(G)]
            assert ranges[0][0] == 0 and ranges[-1][1] == L
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(G - 1))
            sizes = shard_sizes(L, G)
            assert sum(sizes) == L and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        frame_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    spec = S.get_spec("C2")
    g = golden("config_C2")
    ws, bs = golden_weights(g, 3)
    params = [torch.nn.Parameter(t.double().clone()) for pair in zip(ws, bs) for t in pair]
    x = torch.from_numpy(g["x"]).double()
    cot = torch.from_numpy(g["cot"]).double()
    s, e = frame_range(x.shape[0], rank, world)
    y = oracle_model(spec, params[0::2], params[1::2])(x[s:e])
    loss = (y * cot[s:e]).sum()
    loss.backward()
    total = allreduce_flat_grads(params, extra=loss.detach().reshape(1))
    # sharded inference: outputs of the shards concatenate to the single-process result
    outs = [torch.zeros(frame_range(x.shape[0], r, world)[1] - frame_range(x.shape[0], r, world)[0], 2,
                        dtype=torch.float64) for r in range(world)]
    dist.all_gather(outs, y.detach()) if len(set(o.shape for o in outs)) == 1 else None
    if rank == 0:
        ret["grads"] = [p.grad.clone() for p in params]
        ret["loss"] = float(total)
        ret["y"] = torch.cat(outs)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process():
    spec = S.get_spec("C2")
    g = golden("config_C2")
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        grads, loss, y = ret["grads"], ret["loss"], ret["y"]
    names = [("ann_layers.%dth_layer.weight" % k, "ann_layers.%dth_layer.bias" % k) for k in (1, 2, 3)]
    flat = [n for pair in names for n in pair]
    for gp, name in zip(grads, flat):
        ref = torch.from_numpy(g["gp64::" + name])
        assert torch.allclose(gp, ref, rtol=1e-10, atol=1e-12), name
    assert abs(loss - float((torch.from_numpy(g["y64"]) * torch.from_numpy(g["cot"]).double()).sum())) < 1e-9
    assert torch.allclose(y, torch.from_numpy(g["y64"]), atol=1e-12)


class _OracleEncoder(torch.nn.Module):
    """CPU stand-in with the MolANN surface AutoencoderStep uses (the product has no CPU path): the oracle restatement
    as compute, the C2 MLP as parameters."""

    def __init__(self, spec, ws, bs):
        super().__init__()
        self.spec = spec
        self.ws = torch.nn.ParameterList([torch.nn.Parameter(w.double().clone()) for w in ws])
        self.bs = torch.nn.ParameterList([torch.nn.Parameter(b.double().clone()) for b in bs])

    def get_preprocessing_layer(self):
        from helpers import oracle_preprocess
        return oracle_preprocess(self.spec)

    def forward(self, x):
        return oracle_model(self.spec, list(self.ws), list(self.bs))(x)


def _make_c4(spec, g):
    ws, bs = golden_weights(g, 3)
    enc = _OracleEncoder(spec, ws, bs)
    torch.manual_seed(404)
    dec = torch.nn.Sequential(torch.nn.Linear(2, 16), torch.nn.Tanh(), torch.nn.Linear(16, 30)).double()
    return enc, dec


def _train_worker(rank, world, port, ret):
    from molann_b200.train import AutoencoderStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    spec = S.get_spec("C2")
    g = golden("config_C2")
    enc, dec = _make_c4(spec, g)
    x = torch.from_numpy(g["x"]).double()
    s, e = frame_range(x.shape[0], rank, world)
    trainer = AutoencoderStep(enc, dec, lr=1e-2, global_frames=x.shape[0])
    losses = [float(trainer.step(x[s:e])) for _ in range(2)]
    if rank == 0:
        ret["losses"] = losses
        ret["params"] = [p.detach().clone() for p in trainer.params]
    dist.destroy_process_group()


def test_two_rank_autoencoder_step_equals_single_process():
    """C4 host logic: sharded batch + ONE flat allreduce (gradients + loss) + SGD == the single-process step."""
    from molann_b200.train import AutoencoderStep
    spec = S.get_spec("C2")
    g = golden("config_C2")
    enc, dec = _make_c4(spec, g)
    x = torch.from_numpy(g["x"]).double()
    single = AutoencoderStep(enc, dec, lr=1e-2, global_frames=x.shape[0])
    want = [float(single.step(x)) for _ in range(2)]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_train_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        losses, params = ret["losses"], ret["params"]
    assert all(abs(a - b) < 1e-12 * max(1.0, abs(b)) for a, b in zip(losses, want))
    assert want[1] < want[0]
    for p, q in zip(params, single.params):
        assert torch.allclose(p, q.detach(), rtol=1e-10, atol=1e-13)
