"""Shared test utilities: golden loading, oracle drivers, C-ABI plan construction, tolerances."""
import ctypes
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from molann_b200 import plan as P            # noqa: E402
from molann_b200 import synthetic as S       # noqa: E402
from oracle import restatement as R          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TYPE_IDS = {"angle": 0, "bond": 1, "dihedral": 2, "position": 3}


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def spec_program(spec):
    """(local align idx or None, centred ref_x fp32, [(type, local idx)], entries[E,6], d_feat)."""
    inp = list(spec.input_ix)
    feats = [(TYPE_IDS[tp], [inp.index(i) for i in ix]) for (_, tp, ix) in spec.features]
    entries, d = P.compile_feature_program(feats, spec.use_angle_value)
    if spec.align_ix is not None:
        aidx = [inp.index(i) for i in spec.align_ix]
        ref = torch.from_numpy(spec.positions[np.asarray(spec.align_ix)].copy())
        ref = ref - torch.mean(ref, 0)
    else:
        aidx, ref = None, torch.zeros(0, 3)
    return aidx, ref, feats, entries, d


def golden_weights(gold, n_layers):
    ws = [torch.from_numpy(gold["sd::ann_layers.%dth_layer.weight" % (k + 1)]) for k in range(n_layers)]
    bs = [torch.from_numpy(gold["sd::ann_layers.%dth_layer.bias" % (k + 1)]) for k in range(n_layers)]
    return ws, bs


def oracle_model(spec, weights, biases, dtype=torch.float64):
    aidx, ref, feats, _, _ = spec_program(spec)
    ws = [w.to(dtype) for w in weights]
    bs = [b.to(dtype) for b in biases]

    def fn(x):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return R.molann_forward(x.to(dtype), aidx, ref.to(dtype), feats, spec.use_angle_value, ws, bs,
                                    spec.activation)
    return fn


def oracle_preprocess(spec, dtype=torch.float64):
    aidx, ref, feats, _, _ = spec_program(spec)

    def fn(x):
        return R.preprocess_forward(x.to(dtype), aidx, ref.to(dtype), feats, spec.use_angle_value)
    return fn


def oracle_value_and_grad(fn, x, cot, dtype=torch.float64):
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    y = fn(xx)
    (gx,) = torch.autograd.grad((y * cot.to(dtype)).sum(), xx)
    return y.detach(), gx


def frame_rel_err(a, b, floor=0.0):
    """Per-frame max-abs error relative to the per-frame max-abs of the reference value ``b`` (never less than
    ``floor``: for quantities that can pass through zero, e.g. two MLP outputs of a random frame)."""
    a = torch.as_tensor(a).double().reshape(a.shape[0], -1)
    b = torch.as_tensor(b).double().reshape(b.shape[0], -1)
    denom = b.abs().amax(dim=1).clamp_min(max(floor, 1e-30))
    return ((a - b).abs().amax(dim=1) / denom)


def assert_parity(new, ref64, ref32=None, tol=1e-5, what="", floor=0.0):
    """SURVEY 8(c) acceptance: per frame, |new-ref64|/|ref64| <= max(tol, 2 |ref32-ref64|/|ref64|)."""
    err = frame_rel_err(new, ref64, floor)
    bound = torch.full_like(err, tol)
    if ref32 is not None:
        bound = torch.maximum(bound, 2.0 * frame_rel_err(ref32, ref64, floor))
    bad = err > bound
    assert not bool(bad.any()), "%s: %d/%d frames out of tolerance, worst %.3e (bound %.3e)" % (
        what, int(bad.sum()), err.numel(), float(err.max()), float(bound[err.argmax()]))
    return float(err.max())


# ------------------------------------------------------------------------------------------------
# C ABI (ctypes) drivers on torch CUDA tensors
# ------------------------------------------------------------------------------------------------
class CPlan(object):
    """Owns the device arrays of a MolannPlan built from a SystemSpec (+ optional MLP weights)."""

    def __init__(self, spec, weights=None, biases=None, device="cuda"):
        from molann_b200 import _lib
        self.lib = _lib.cabi()
        aidx, ref, feats, entries, d = spec_program(spec)
        dev = torch.device(device)
        self.keep = []
        p = _lib.MolannPlan()
        p.n_inp = spec.n_inp
        if aidx is not None:
            a = torch.tensor(aidx, dtype=torch.int32, device=dev)
            r = ref.to(dev).contiguous()
            self.keep += [a, r]
            p.n_align, p.align_idx, p.ref_x = len(aidx), a.data_ptr(), r.data_ptr()
        e = torch.from_numpy(entries).to(dev).contiguous()
        self.keep.append(e)
        p.n_entries, p.entries, p.d_feat = entries.shape[0], e.data_ptr(), d
        p.use_angle_value = int(spec.use_angle_value)
        act_ids = {"tanh": 0, "relu": 1, "sigmoid": 2, "identity": 3}
        if weights is not None:
            p.n_layers, p.act_id = len(weights), act_ids[spec.activation]
            p.dims[0] = d
            for k, (w, b) in enumerate(zip(weights, biases)):
                wd, bd = w.to(dev).float().contiguous(), b.to(dev).float().contiguous()
                self.keep += [wd, bd]
                p.W[k], p.b[k] = wd.data_ptr(), bd.data_ptr()
                p.dims[k + 1] = w.shape[0]
        self.p = p
        self.spec = spec
        self.device = dev

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ws(self, L, backward):
        n = self.lib.molann_b200_workspace_bytes(ctypes.byref(self.p), L, int(backward))
        return torch.empty(max(n, 1), dtype=torch.uint8, device=self.device), n

    def forward(self, x):
        from molann_b200 import _lib
        L = x.shape[0]
        y = torch.empty(L, self.p.dims[self.p.n_layers], device=self.device)
        ws, n = self._ws(L, False)
        _lib.check(self.lib.molann_b200_forward(ctypes.byref(self.p), x.data_ptr(), L, y.data_ptr(), ws.data_ptr(), n,
                                                self._stream()), "forward")
        return y

    def prepare(self):
        """Build the prepared plan (include/molann_b200.h) once; returns the opaque handle."""
        from molann_b200 import _lib
        if getattr(self, "_prep", None):
            return self._prep
        n = self.lib.molann_b200_prepared_bytes(ctypes.byref(self.p))
        assert n > 0, "plan not eligible for the prepared (wide) path"
        self._prep_buf = torch.empty(n, dtype=torch.uint8, device=self.device)
        h = ctypes.c_void_p()
        _lib.check(self.lib.molann_b200_prepare(ctypes.byref(self.p), self._prep_buf.data_ptr(), n, self._stream(),
                                                ctypes.byref(h)), "prepare")
        self._prep = h
        return h

    def forward_prepared(self, x):
        from molann_b200 import _lib
        h = self.prepare()
        L = x.shape[0]
        y = torch.empty(L, self.p.dims[self.p.n_layers], device=self.device)
        n = self.lib.molann_b200_prepared_workspace_bytes(h, L)
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.molann_b200_forward_prepared(h, ctypes.byref(self.p), x.data_ptr(), L, y.data_ptr(),
                                                         ws.data_ptr(), n, self._stream()), "forward_prepared")
        return y

    def refresh(self):
        from molann_b200 import _lib
        _lib.check(self.lib.molann_b200_prepared_refresh(self.prepare(), ctypes.byref(self.p), self._stream()),
                   "prepared_refresh")

    def __del__(self):
        if getattr(self, "_prep", None):
            self.lib.molann_b200_prepared_destroy(self._prep)
            self._prep = None

    def backward(self, x, gy, want_params=False):
        from molann_b200 import _lib
        L = x.shape[0]
        gx = torch.empty_like(x)
        ws, n = self._ws(L, True)
        gW = gb = None
        outs = []
        if want_params:
            nl = self.p.n_layers
            gWt = [torch.zeros(self.p.dims[k + 1], self.p.dims[k], device=self.device) for k in range(nl)]
            gbt = [torch.zeros(self.p.dims[k + 1], device=self.device) for k in range(nl)]
            gW = (ctypes.c_void_p * nl)(*[t.data_ptr() for t in gWt])
            gb = (ctypes.c_void_p * nl)(*[t.data_ptr() for t in gbt])
            outs = [gWt, gbt]
        _lib.check(self.lib.molann_b200_backward(ctypes.byref(self.p), x.data_ptr(), gy.contiguous().data_ptr(), L,
                                                 gx.data_ptr(), gW, gb, ws.data_ptr(), n, self._stream()), "backward")
        return (gx, *outs) if want_params else gx

    def preprocess_forward(self, x):
        from molann_b200 import _lib
        L = x.shape[0]
        f = torch.empty(L, self.p.d_feat, device=self.device)
        _lib.check(self.lib.molann_b200_preprocess_forward(ctypes.byref(self.p), x.data_ptr(), L, f.data_ptr(),
                                                           self._stream()), "preprocess_forward")
        return f

    def preprocess_backward(self, x, gfeat):
        from molann_b200 import _lib
        L = x.shape[0]
        gx = torch.empty_like(x)
        _lib.check(self.lib.molann_b200_preprocess_backward(ctypes.byref(self.p), x.data_ptr(),
                                                            gfeat.contiguous().data_ptr(), L, gx.data_ptr(),
                                                            self._stream()), "preprocess_backward")
        return gx

    def align_forward(self, x):
        from molann_b200 import _lib
        out = torch.empty_like(x)
        _lib.check(self.lib.molann_b200_align_forward(ctypes.byref(self.p), x.data_ptr(), x.shape[0], out.data_ptr(),
                                                      self._stream()), "align_forward")
        return out

    def align_backward(self, x, gout):
        from molann_b200 import _lib
        gx = torch.empty_like(x)
        _lib.check(self.lib.molann_b200_align_backward(ctypes.byref(self.p), x.data_ptr(), gout.contiguous().data_ptr(),
                                                       x.shape[0], gx.data_ptr(), self._stream()), "align_backward")
        return gx
