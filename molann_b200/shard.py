"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

Frames are independent (no reference op reduces over the frame axis: molann/ann.py:179-197,323-354),
so inference and d/dx shard as contiguous frame ranges with NO inter-GPU traffic.  The only collective
on the path is the sum-allreduce of the flat MLP-parameter gradient in training (NCCL on GPUs; the
same code runs over gloo on CPU in the tests).
"""
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard ``[rank*L/G, (rank+1)*L/G)`` (integer arithmetic, covers all frames exactly once)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    return (n_frames * rank) // world_size, (n_frames * (rank + 1)) // world_size


def shard_sizes(n_frames: int, world_size: int) -> List[int]:
    return [frame_range(n_frames, r, world_size)[1] - frame_range(n_frames, r, world_size)[0]
            for r in range(world_size)]


def allreduce_flat_grads(params: Iterable[torch.nn.Parameter], average: bool = False,
                         group: Optional[dist.ProcessGroup] = None, extra: Optional[torch.Tensor] = None):
    """ONE sum-allreduce over a flat buffer holding every parameter gradient (+ optional scalars such as
    the loss in ``extra``), then scatter back into ``p.grad``.  Returns the reduced ``extra`` (or None).

    Message size is tens of KB for the molann MLPs, so the collective is latency-bound: one launch.
    """
    params = [p for p in params if p.grad is not None]
    if not params:
        return extra
    chunks = [p.grad.reshape(-1) for p in params]
    if extra is not None:
        chunks.append(extra.reshape(-1).to(chunks[0].dtype))
    flat = torch.cat(chunks)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    if extra is not None:
        return flat[off:off + extra.numel()].view_as(extra)
    return None
