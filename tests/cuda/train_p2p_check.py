"""Two or more ranks (torchrun): the fused training step's collective over NVLink peer memory
(molann_b200_allreduce_sgd: one-shot allreduce in rank order + SGD in one kernel) against (a) the full-batch gradient
computed on one GPU and (b) the NCCL route of the same step; replicas must stay bit-identical; CUDA-graph replay."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build():
    from molann_b200 import synthetic as S
    spec = S.get_spec("C2")
    enc, _ = S.build_model(spec)
    torch.manual_seed(11)
    dec = S.default_api().create_sequential_nn([spec.out_dim(), 64, 64, spec.feature_dim()])
    return spec, enc.cuda(), dec.cuda()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    from molann_b200 import synthetic as S
    from molann_b200.shard import frame_range
    from molann_b200.train import AutoencoderStep
    L = 4096 + 77
    spec, enc, dec = build()
    x_full = S.make_frames(spec, L, seed=5).cuda()
    lo, hi = frame_range(L, rank, world)
    x = x_full[lo:hi].contiguous()

    def run(p2p, steps):
        os.environ["MOLANN_B200_TRAIN_P2P"] = "1" if p2p else "0"
        _, e, d = build()
        tr = AutoencoderStep(e, d, lr=0.05, global_frames=L)
        first = tr.loss_and_grads(x).clone()
        flat1 = tr._flat.clone()
        losses = [float(tr.step(x)) for _ in range(steps)]
        return tr, first, flat1, losses

    tr_a, first_a, flat_a, losses_a = run(True, 3)
    assert tr_a._fused_args is not None, "fused training kernel not selected"
    assert tr_a._peer is not None, "peer-memory route not taken: %s" % tr_a._peer_error
    tr_b, first_b, flat_b, losses_b = run(False, 3)
    assert tr_b._peer is None
    # (a) the collective's sum is the full-batch gradient
    _, e0, d0 = build()
    ref = AutoencoderStep(e0, d0, lr=0.05, global_frames=L)
    args = ref._fused_call_args(x_full)
    geo0, geo1, entries, d_feat, use_angle, enc_params, enc_act, dec_params, dec_act = args
    full = torch.ops.molann_b200.train_loss_and_grads(x_full, geo0, geo1, entries, d_feat, use_angle,
                                                      [q.detach() for q in enc_params], enc_act,
                                                      [q.detach() for q in dec_params], dec_act, 1.0 / (L * d_feat))
    scale = float(full[:-1].abs().max())
    assert float((flat_a - full).abs().max()) < 2e-6 * scale + 1e-7 * abs(float(full[-1])), "p2p sum != full batch"
    assert float((flat_b - full).abs().max()) < 2e-6 * scale + 1e-7 * abs(float(full[-1])), "nccl sum != full batch"
    # (b) both routes train alike (their summation orders differ) and the loss falls
    for a, b in zip(losses_a, losses_b):
        assert abs(a - b) < 1e-5 * abs(b), (losses_a, losses_b)
    assert losses_a[-1] < float(first_a)
    for p, q in zip(tr_a.params, tr_b.params):
        assert float((p.detach() - q.detach()).abs().max()) < 1e-5 * float(q.detach().abs().max())
    # replicas are bit-identical
    mine = torch.cat([p.detach().reshape(-1) for p in tr_a.params])
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "replicas drifted"
    # the step as a CUDA graph
    assert tr_a.capture(x), tr_a._capture_error
    l1 = float(tr_a.replay())
    l2 = float(tr_a.replay())
    assert l2 < l1 < losses_a[-1] * 1.0001, (losses_a, l1, l2)
    mine = torch.cat([p.detach().reshape(-1) for p in tr_a.params])
    dist.all_gather(gathered, mine)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "replicas drifted under graph replay"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("P2P_OK world=%d losses=%s graph=%s" % (world, ["%.6f" % v for v in losses_a], ["%.6f" % l1, "%.6f" % l2]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
