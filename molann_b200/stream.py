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


def quantize_frames(x: torch.Tensor, resolution: Optional[float] = None):
    """Encode fp32 frames ``[L, n, 3]`` (host) into the int16 wire format: ``x ~ origin + q * resolution``.

    ``origin`` is the centre of the batch's bounding box per axis; ``resolution`` defaults to the finest step that
    covers the box (range / 65534), and is never finer than what the caller asks for.  Returns
    ``(q int16 [L, n, 3], origin (3 floats), resolution)``.  The coding error is at most resolution / 2 per coordinate
    (XTC files store 0.01 Angstrom steps); it is a property of the transport -- results are exact functions of the
    DECODED coordinates, see :func:`dequantize_frames`."""
    xx = x.detach().to("cpu", torch.float32)
    lo = xx.amin(dim=(0, 1)).double()
    hi = xx.amax(dim=(0, 1)).double()
    origin = ((lo + hi) * 0.5).float()
    need = float(((hi - lo).max() / 65534.0).item())
    res = max(need * (1.0 + 1e-6), float(resolution) if resolution is not None else 0.0, 1e-30)
    q = torch.round((xx.double() - origin.double()) / res).clamp_(-32767, 32767).to(torch.int16)
    return q.contiguous(), [float(v) for v in origin.tolist()], float(res)


def dequantize_frames(q: torch.Tensor, origin, resolution: float) -> torch.Tensor:
    """Host decoder, bit-identical to the device one: ONE fp32 fused multiply-add per coordinate, done here as the
    correctly rounded fp32 value of the exact fp64 expression (an int16 times an fp32 is exact in fp64)."""
    o = torch.tensor([float(v) for v in origin], dtype=torch.float32).double()
    r = float(torch.tensor(resolution, dtype=torch.float32))
    return (q.double() * r + o).float()


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
        self.qbuf = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run_wire(self, q_host: torch.Tensor, origin, resolution: float, y_host: torch.Tensor):
        """``y_host[:] = model(decode(q_host))`` for int16 wire frames (:func:`quantize_frames`): half the H2D bytes of
        :meth:`run`; the decode is one small kernel per chunk on the device (2 % of the chunk's PCIe time)."""
        assert q_host.is_pinned() and y_host.is_pinned(), "host buffers must be pinned"
        assert q_host.dtype == torch.int16 and q_host.dim() == 3 and q_host.shape[1] == self.n_inp
        L = q_host.shape[0]
        if self.qbuf is None:
            self.qbuf = [torch.empty(self.chunk, self.n_inp, 3, device=self.device, dtype=torch.int16) for _ in range(2)]
        main = torch.cuda.current_stream(self.device)
        in_done = [torch.cuda.Event() for _ in range(2)]
        comp_done = [torch.cuda.Event() for _ in range(2)]
        out_done = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = self.d2h_bytes = 0
        ox, oy, oz = [float(v) for v in origin]
        for c in range((L + self.chunk - 1) // self.chunk):
            s, e = c * self.chunk, min(L, (c + 1) * self.chunk)
            m, b = e - s, c & 1
            with torch.cuda.stream(self.copy_in):
                if c >= 2:
                    self.copy_in.wait_event(comp_done[b])
                self.qbuf[b][:m].copy_(q_host[s:e], non_blocking=True)
                self.h2d_bytes += m * self.n_inp * 3 * 2
                in_done[b].record(self.copy_in)
            main.wait_event(in_done[b])
            if c >= 2:
                main.wait_event(out_done[b])
            with torch.no_grad():
                xb = torch.ops.molann_b200.decode_frames(self.qbuf[b][:m], ox, oy, oz, float(resolution))
                self.ybuf[b][:m].copy_(self.model(xb))
            comp_done[b].record(main)
            with torch.cuda.stream(self.copy_out), torch.no_grad():
                self.copy_out.wait_event(comp_done[b])
                y_host[s:e].copy_(self.ybuf[b][:m], non_blocking=True)
                self.d2h_bytes += m * self.out_dim * 4
                out_done[b].record(self.copy_out)
        main.wait_stream(self.copy_out)
        done = torch.cuda.Event()
        done.record(self.copy_out)
        done.synchronize()
        return y_host

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
