"""Host-buffer streaming front end: pinned host frames -> H2D -> fused kernels -> D2H results.

This is the end-to-end entry a trajectory reader uses (the reference only ever shows
``torch.tensor(ag.positions).unsqueeze(0)``, molann/ann.py:106): frames live in (pinned) host memory,
are copied chunk by chunk on a copy stream while the previous chunk computes, and outputs (and,
optionally, coordinate gradients) are copied back.  PyTorch supplies streams/events/pinned memory only.
"""
from typing import Optional

import torch


def nvml_handle(pynvml, device_index: int):
    """NVML handle of CUDA device ``device_index`` (by UUID: CUDA_VISIBLE_DEVICES may renumber the devices)."""
    try:
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        return pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
    except Exception:
        return pynvml.nvmlDeviceGetHandleByIndex(device_index)


def bind_to_gpu_numa_node(device_index: int) -> int:
    """Restrict the calling process to the CPUs NVML reports as closest to GPU ``device_index`` (same NUMA node /
    PCIe root).  Pinned buffers allocated afterwards are first-touched on that node, so H2D / D2H copies of an
    8-GPU box do not cross the inter-socket link.  Returns the number of CPUs kept (0: left unchanged)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = nvml_handle(pynvml, device_index)
        words = (max(os.sched_getaffinity(0)) // 64) + 1
        try:
            ncpu = os.cpu_count() or 64
            words = max(words, (ncpu + 63) // 64)
        except Exception:
            pass
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        ideal = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        keep = ideal & os.sched_getaffinity(0)
        if keep:
            os.sched_setaffinity(0, keep)
            return len(keep)
    except Exception:
        pass
    return 0


class HostPipeline(object):
    """Double-buffered host->device->host pipeline around a ``molann_b200`` model.

    Args:
        model: a CUDA-resident module from :mod:`molann_b200.ann` (``MolANN`` or ``PreprocessingANN``).
        n_inp, out_dim: atoms per frame / model outputs per frame.
        chunk_frames: frames per H2D chunk.

    ``run`` returns only after the last device->host copy has landed (it synchronises on the copy-out stream), so
    the caller may read ``y_host`` / ``gx_host`` and reuse ``x_host`` immediately.  Nothing it writes carries
    autograd history.
    """

    def __init__(self, model, n_inp: int, out_dim: int, chunk_frames: int = 1 << 18, device=None):
        self.model = model
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.chunk = int(chunk_frames)
        self.n_inp, self.out_dim = int(n_inp), int(out_dim)
        self.copy_in = torch.cuda.Stream(device=self.device)
        self.copy_out = torch.cuda.Stream(device=self.device)
        self.xbuf = [torch.empty(self.chunk, n_inp, 3, device=self.device) for _ in range(2)]
        self.ybuf = [torch.empty(self.chunk, out_dim, device=self.device) for _ in range(2)]
        self.gbuf = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, x_host: torch.Tensor, y_host: torch.Tensor, cot_host: Optional[torch.Tensor] = None,
            gx_host: Optional[torch.Tensor] = None):
        """``y_host[:] = model(x_host)``; with ``cot_host``/``gx_host`` also ``gx_host[:] = d<cot,y>/dx``.
        Blocks the host until the results are in the host buffers."""
        assert x_host.is_pinned() and y_host.is_pinned(), "host buffers must be pinned"
        assert not y_host.requires_grad and (gx_host is None or not gx_host.requires_grad), \
            "result buffers must be plain tensors"
        L = x_host.shape[0]
        want_grad = gx_host is not None
        if want_grad and self.gbuf is None:
            self.gbuf = [torch.empty(self.chunk, self.out_dim, device=self.device) for _ in range(2)]
        main = torch.cuda.current_stream(self.device)
        in_done = [torch.cuda.Event() for _ in range(2)]
        comp_done = [torch.cuda.Event() for _ in range(2)]
        out_done = [torch.cuda.Event() for _ in range(2)]
        keep = [None, None]
        self.h2d_bytes = self.d2h_bytes = 0
        nchunks = (L + self.chunk - 1) // self.chunk
        for c in range(nchunks):
            s, e = c * self.chunk, min(L, (c + 1) * self.chunk)
            m, b = e - s, c & 1
            with torch.cuda.stream(self.copy_in):
                if c >= 2:
                    self.copy_in.wait_event(comp_done[b])       # buffer b free again
                self.xbuf[b][:m].copy_(x_host[s:e], non_blocking=True)
                self.h2d_bytes += x_host[s:e].numel() * 4
                if want_grad:
                    self.gbuf[b][:m].copy_(cot_host[s:e], non_blocking=True)
                    self.h2d_bytes += cot_host[s:e].numel() * 4
                in_done[b].record(self.copy_in)
            main.wait_event(in_done[b])
            if c >= 2:
                main.wait_event(out_done[b])                    # ybuf[b] drained
            xb = self.xbuf[b][:m]
            if want_grad:
                if hasattr(self.model, "value_and_grad"):
                    y, gx = self.model.value_and_grad(xb, self.gbuf[b][:m])
                else:
                    xb = xb.detach().requires_grad_(True)
                    y = self.model(xb)
                    (gx,) = torch.autograd.grad(y, xb, self.gbuf[b][:m])
                with torch.no_grad():
                    self.ybuf[b][:m].copy_(y.detach())
                keep[b] = gx.detach()
            else:
                with torch.no_grad():
                    self.ybuf[b][:m].copy_(self.model(xb))
            comp_done[b].record(main)
            with torch.cuda.stream(self.copy_out), torch.no_grad():
                self.copy_out.wait_event(comp_done[b])
                y_host[s:e].copy_(self.ybuf[b][:m], non_blocking=True)
                self.d2h_bytes += m * self.out_dim * 4
                if want_grad:
                    gx_host[s:e].copy_(keep[b], non_blocking=True)
                    keep[b].record_stream(self.copy_out)
                    self.d2h_bytes += m * self.n_inp * 12
                out_done[b].record(self.copy_out)
        main.wait_stream(self.copy_out)
        done = torch.cuda.Event()
        done.record(self.copy_out)
        done.synchronize()                  # the D2H copies (and every H2D read of x_host before them) have finished
        return y_host
