"""Data-parallel autoencoder training step on top of the fused kernels (BASELINE.json configs[3] / SURVEY 8(d) C4).

The reference ships no training loop; C4 is an *external* use of its modules: encoder = ``MolANN``
(molann/ann.py:567-624), decoder = ``create_sequential_nn`` (ann.py:37-67), target = the encoder's own
``PreprocessingANN`` output (ann.py:553-565), loss = mean squared reconstruction error.  Every rank owns a
contiguous shard of the global batch (molann_b200.shard.frame_range); the ONLY collective is one flat
sum-allreduce of all MLP parameter gradients (+ the loss scalar) per step (shard.allreduce_flat_grads: NCCL on
GPUs, gloo in the CPU tests).  The encoder forward/backward run through ``molann_b200::molann`` (fused align +
features + MLP kernels, parameter gradients from the C ABI's ``molann_b200_backward``); the decoder is a plain
``torch.nn.Sequential`` (library GEMMs).
"""
from typing import Optional

import torch

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

    def loss_and_grads(self, x_local: torch.Tensor) -> torch.Tensor:
        """Forward + backward on this rank's shard, then the single flat allreduce.  Returns the global loss."""
        n_global = self.global_frames if self.global_frames is not None else x_local.shape[0]
        for p in self.params:
            p.grad = None
        with torch.no_grad():
            target = self.encoder.get_preprocessing_layer()(x_local)
        recon = self.decoder(self.encoder(x_local))
        loss = ((recon - target) ** 2).sum() / float(n_global * target.shape[1])
        loss.backward()
        total = allreduce_flat_grads(self.params, group=self.group, extra=loss.detach().reshape(1))
        return total if total is not None else loss.detach().reshape(1)

    def step(self, x_local: torch.Tensor) -> torch.Tensor:
        loss = self.loss_and_grads(x_local)
        with torch.no_grad():
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
