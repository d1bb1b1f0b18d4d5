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
