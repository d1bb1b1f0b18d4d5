"""Data-parallel autoencoder training step on top of the fused kernels (BASELINE.json configs[3] / SURVEY 8(d) C4).

The reference ships no training loop; C4 is an *external* use of its modules: encoder = ``MolANN``
(molann/ann.py:567-624), decoder = ``create_sequential_nn`` (ann.py:37-67), target = the encoder's own
``PreprocessingANN`` output (ann.py:553-565), loss = mean squared reconstruction error.  Every rank owns a
contiguous shard of the global batch (molann_b200.shard.frame_range); the ONLY collective is one flat
sum-allreduce of all MLP parameter gradients (+ the loss scalar) per step (shard.allreduce_flat_grads: NCCL on
GPUs, gloo in the CPU tests).  The encoder forward/backward run through ``molann_b200::molann`` (fused align +
features + MLP kernels, parameter gradients from the C ABI's ``molann_b200_backward``); the decoder is a plain
``torch.nn.Sequential`` (library GEMMs).

When encoder and decoder are both ``create_sequential_nn``-style stacks small enough for the SM (C4 is), the whole
step is the fused training kernel instead (``molann_b200::train_loss_and_grads``, csrc/fused_train.cuh): forward of
encoder and decoder, loss, every parameter gradient and the loss in ONE kernel plus a fixed-order reduction, and the
SGD update in one more launch (``molann_b200::sgd_apply_``) -- no library GEMM, no autograd graph.  With several ranks
the flat result is summed over NVLink peer memory and applied by ONE kernel per rank (``molann_b200::allreduce_sgd_``:
one-shot allreduce in rank order, bit-identical replicas, no NCCL call; ``MOLANN_B200_TRAIN_P2P=0`` uses NCCL +
``sgd_apply_``).  ``step_from_host`` streams a pinned host shard in pieces under the kernel.
``MOLANN_B200_TRAIN_FUSED=0`` keeps the composed path.
"""
import os
from typing import Optional

import torch

import torch.distributed as dist

from .shard import allreduce_flat_grads


class AutoencoderStep(object):
    """One SGD step of ``loss = mean((decoder(encoder(x)) - preprocessing(x))**2)`` over the GLOBAL batch.

    Args:
        encoder: a ``MolANN`` (any device); its preprocessing layer provides the reconstruction target.
        decoder: module mapping ``[l, k] -> [l, d_feat]``.
        lr: plain SGD learning rate (1e-3).
        global_frames: frames in the global batch (the loss is normalised by it on every rank, so the
            allreduced gradient equals the single-process gradient of the concatenated batch).
    """

    def __init__(self, encoder, decoder, lr: float = 1e-3, global_frames: Optional[int] = None, group=None):
        self.encoder, self.decoder = encoder, decoder
        self.params = [p for p in list(encoder.parameters()) + list(decoder.parameters()) if p.requires_grad]
        self.lr = float(lr)
        self.global_frames = global_frames
        self.group = group
        self._fused_args = None
        self._fused_checked = False
        self._peer = None
        self._peer_checked = False
        self._peer_error = None
        self._updated = False

    # ---- the fused training kernel ---------------------------------------------------------------------------------
    def _fused_call_args(self, x_local):
        """Arguments of ``molann_b200::train_loss_and_grads`` when the fused kernel serves this pair, else None
        (decided on the first CUDA batch: the modules must already live on the device)."""
        if self._fused_checked:
            return self._fused_args
        self._fused_checked = True
        if not x_local.is_cuda or os.environ.get("MOLANN_B200_TRAIN_FUSED", "1") == "0":
            return None
        from . import ann as _ann
        enc, dec = self.encoder, self.decoder
        if not isinstance(enc, _ann.MolANN) or not enc._fused or not enc._mlp_unchanged():
            return None
        dec_act = _ann._fusable_activation(dec)
        if dec_act < 0:
            return None
        enc_params = [t for layer in enc.ann_layers if hasattr(layer, 'weight') for t in (layer.weight, layer.bias)]
        dec_params = [t for layer in dec if hasattr(layer, 'weight') for t in (layer.weight, layer.bias)]
        ordered = enc_params + dec_params
        if len(ordered) != len(self.params) or any(a is not b for a, b in zip(ordered, self.params)):
            return None                       # frozen or extra parameters: the flat layout would not match
        if any(not q.is_contiguous() or q.dtype != torch.float32 or q.device != x_local.device for q in ordered):
            return None
        pp, flayer = enc.preprocessing_layer, enc.preprocessing_layer.feature_layer
        if enc._fused_align:
            geo = (pp.align_layer._align_idx, pp.align_layer.ref_x)
        else:
            geo = (flayer._no_idx, flayer._no_ref)
        args = geo + (flayer._entries, flayer._dim, flayer.use_angle_value, enc_params, enc._act_id, dec_params, dec_act)
        with torch.no_grad():
            if not torch.ops.molann_b200.train_eligible(x_local, *[[q.detach() for q in a] if isinstance(a, list) else a
                                                                   for a in args]):
                return None
        self._fused_args = args
        return args

    # ---- the collective of the fused path: one-shot allreduce over NVLink peer memory, fused with the SGD update ------
    def _peer_setup(self, flat):
        """Symmetric buffers of all ranks of the group (torch symmetric memory: CUDA IPC / fabric handles), or None --
        then the flat vector goes through NCCL.  ``MOLANN_B200_TRAIN_P2P=0`` forces NCCL."""
        if self._peer_checked:
            return self._peer
        self._peer_checked = True
        self._peer = None
        if os.environ.get("MOLANN_B200_TRAIN_P2P", "1") == "0":
            return None
        world = dist.get_world_size(self.group)
        ok = torch.ones(1, device=flat.device)
        try:
            if world > 8 or "nccl" not in str(dist.get_backend(self.group)):
                raise RuntimeError("one node, NCCL group, at most 8 ranks")
            import torch.distributed._symmetric_memory as symm
            nbytes = torch.ops.molann_b200.allreduce_buffer_bytes(flat.numel(), world)
            buf = symm.empty(nbytes // 4, dtype=torch.float32, device=flat.device)
            buf.zero_()
            hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            ptrs = [int(a) for a in hdl.buffer_ptrs]
            state = torch.tensor([1, 0, 0], dtype=torch.int32, device=flat.device)
            peer = {"buf": buf, "hdl": hdl, "ptrs": ptrs, "rank": dist.get_rank(self.group), "state": state}
        except Exception as exc:  # noqa: BLE001 -- no peer access on this stack: NCCL carries the vector
            self._peer_error = repr(exc)
            peer = None
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # all ranks take the same route
        torch.cuda.synchronize()                                        # every buffer is zeroed before anyone signals
        dist.barrier(group=self.group)
        self._peer = peer if bool(ok.item() > 0.5) else None
        return self._peer

    def _fused_loss_and_grads(self, x_local, args, n_global, lr=0.0):
        geo0, geo1, entries, d_feat, use_angle, enc_params, enc_act, dec_params, dec_act = args
        self._updated = False
        with torch.no_grad():
            flat = torch.ops.molann_b200.train_loss_and_grads(
                x_local, geo0, geo1, entries, d_feat, use_angle, [q.detach() for q in enc_params], enc_act,
                [q.detach() for q in dec_params], dec_act, 1.0 / (float(n_global) * float(d_feat)))
        return self._finish_flat(flat, lr)

    def _finish_flat(self, flat, lr):
        """The collective over the ranks' flat vectors (+ the SGD update when the peer-memory kernel runs and lr != 0);
        every ``p.grad`` becomes a view of the global vector.  Returns the global loss."""
        with torch.no_grad():
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                peer = self._peer_setup(flat)
                if peer is not None:          # sum over ranks (in rank order) and, when lr != 0, p -= lr * g: one kernel
                    flat = torch.ops.molann_b200.allreduce_sgd_([q.detach() for q in self.params], flat, peer["ptrs"],
                                                                peer["rank"], peer["state"], lr)
                    self._updated = lr != 0.0
                else:
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            off = 0
            for q in self.params:               # gradients are views of the flat buffer: nothing is copied
                n = q.numel()
                q.grad = flat[off:off + n].view_as(q)
                off += n
        self._flat = flat
        return flat[off:off + 1]

    def loss_and_grads(self, x_local: torch.Tensor, _lr: float = 0.0) -> torch.Tensor:
        """Forward + backward on this rank's shard, then the single flat allreduce.  Returns the global loss."""
        n_global = self.global_frames if self.global_frames is not None else x_local.shape[0]
        self._flat = None
        self._updated = False
        fused = self._fused_call_args(x_local)
        if fused is not None:
            return self._fused_loss_and_grads(x_local, fused, n_global, _lr)
        for p in self.params:
            p.grad = None
        with torch.no_grad():
            target = self.encoder.get_preprocessing_layer()(x_local)
        recon = self.decoder(self.encoder(x_local))
        loss = ((recon - target) ** 2).sum() / float(n_global * target.shape[1])
        loss.backward()
        total = allreduce_flat_grads(self.params, group=self.group, extra=loss.detach().reshape(1))
        return total if total is not None else loss.detach().reshape(1)

    def step_from_host(self, x_host: torch.Tensor, chunks: int = 8) -> torch.Tensor:
        """One SGD step on a shard that lives in (pinned) HOST memory, fused path only: the shard streams to the device
        in ``chunks`` pieces on a copy stream while the fused training kernel works on the previous piece -- the flat
        vectors of the pieces simply add up, because ``loss_scale`` is that of the global batch -- then the collective
        and the update as in :meth:`step`.  Returns the global loss (device tensor)."""
        L = x_host.shape[0]
        n_global = self.global_frames if self.global_frames is not None else L
        dev = self.params[0].device
        per = max(128, -(-L // max(1, chunks)))
        per = -(-per // 128) * 128                              # whole 128-frame tiles per piece
        if getattr(self, "_stage", None) is None or self._stage[0].shape[0] < per or self._stage[0].shape[1:] != x_host.shape[1:]:
            self._stage = [torch.empty((per,) + tuple(x_host.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
        main, copy = torch.cuda.current_stream(dev), self._copy_stream
        fused = self._fused_call_args(self._stage[0])
        if fused is None:
            raise RuntimeError("molann_b200: step_from_host needs the fused training kernel (see train_eligible); "
                               "copy the shard to the device and call step() instead")
        geo0, geo1, entries, d_feat, use_angle, enc_params, enc_act, dec_params, dec_act = fused
        scale = 1.0 / (float(n_global) * float(d_feat))
        self._flat, self._updated = None, False
        free = [None, None]
        flat = None
        copy.wait_stream(main)
        with torch.no_grad():
            for i, s in enumerate(range(0, L, per)):
                e = min(L, s + per)
                buf = self._stage[i % 2]
                with torch.cuda.stream(copy):
                    if free[i % 2] is not None:
                        copy.wait_event(free[i % 2])            # the kernel that read this buffer two pieces ago is done
                    buf[:e - s].copy_(x_host[s:e], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(copy)
                main.wait_event(ready)
                part = torch.ops.molann_b200.train_loss_and_grads(
                    buf[:e - s], geo0, geo1, entries, d_feat, use_angle, [q.detach() for q in enc_params], enc_act,
                    [q.detach() for q in dec_params], dec_act, scale)
                free[i % 2] = torch.cuda.Event()
                free[i % 2].record(main)
                flat = part if flat is None else flat.add_(part)
            if flat is None:
                flat = torch.zeros(sum(q.numel() for q in self.params) + 1, device=dev)
        loss = self._finish_flat(flat, self.lr)
        with torch.no_grad():
            if not self._updated:
                torch.ops.molann_b200.sgd_apply_([q.detach() for q in self.params], self._flat, self.lr)
        return loss

    def step(self, x_local: torch.Tensor) -> torch.Tensor:
        loss = self.loss_and_grads(x_local, self.lr)
        with torch.no_grad():
            if self._updated:                   # the allreduce kernel already applied p -= lr * g
                return loss
            if self._flat is not None:          # fused path: one launch over the flat gradient
                torch.ops.molann_b200.sgd_apply_([q.detach() for q in self.params], self._flat, self.lr)
                return loss
            for p in self.params:
                if p.grad is not None:
                    p.add_(p.grad, alpha=-self.lr)
        return loss

    # ---- the same step as ONE CUDA graph launch ------------------------------------------------------------------
    # A step is ~70 small launches of ours plus the decoder's library kernels; issued one by one the host is the
    # bottleneck (7.9 ms per 2^20-frame step of which ~5 ms is GPU work).  The shapes never change, so the whole step
    # -- forward, backward, the NCCL allreduce and the SGD update -- is captured once and replayed.
    def capture(self, x_static: torch.Tensor, warmup: int = 3):
        """Capture ``step(x_static)``.  Later ``replay()`` calls re-run it on whatever ``x_static`` then holds
        (copy the next shard into it).  Returns False (and stays eager) if the capture is refused."""
        self._graph = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self.step(x_static)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            for p in self.params:
                p.grad = None
            with torch.cuda.graph(graph):
                self._static_loss = self.step(x_static)
            self._graph, self._x_static = graph, x_static
            return True
        except Exception as exc:  # noqa: BLE001 -- e.g. a collective that cannot be captured on this stack
            self._graph = None
            self._capture_error = repr(exc)
            torch.cuda.synchronize()
            return False

    def replay(self) -> torch.Tensor:
        """One captured step on the current contents of the static input; returns the (static) loss tensor."""
        if getattr(self, "_graph", None) is None:
            return self.step(self._x_static)
        self._graph.replay()
        return self._static_loss
