"""Dev probe: C3 parity margins and run-to-run determinism of the general path (GPU)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from molann_b200 import synthetic as S
from helpers import oracle_model, oracle_value_and_grad, frame_rel_err

spec = S.get_spec("C3")
model, _ = S.build_model(spec)
sd = model.state_dict()
ws = [sd["ann_layers.%dth_layer.weight" % k] for k in (1, 2, 3)]
bs = [sd["ann_layers.%dth_layer.bias" % k] for k in (1, 2, 3)]
L = 203
x = S.make_frames(spec, L, seed=17)
cot = torch.randn(L, 2, generator=torch.Generator().manual_seed(5))
y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), x, cot)
y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x, cot, torch.float32)
model = model.cuda()
e32 = frame_rel_err(gx32, gx64)
print("ref32 gx err: max %.3e" % float(e32.max()))
import time
xb = S.make_frames(spec, 32768, device="cuda", seed=3)
cotb = torch.zeros(32768, 2, device="cuda"); cotb[:, 0] = 1
for mode, seg in (("ffma", 0), ("tc", 4), ("tc", 2), ("tc", 1)):
    os.environ["MOLANN_B200_GEMM_TC"] = "1" if mode == "tc" else "0"
    os.environ["MOLANN_B200_GEMM_SEG"] = str(max(seg, 1))
    xd = x.cuda().requires_grad_(True)
    y = model(xd)
    (gx,) = torch.autograd.grad(y, xd, cot.cuda())
    ey = frame_rel_err(y.detach().cpu(), y64); eg = frame_rel_err(gx.cpu(), gx64)
    w = int(eg.argmax())
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    with torch.no_grad():
        tf = t(lambda: model(xb))
    tv = t(lambda: model.value_and_grad(xb, cotb))
    print("%-5s seg %d: y max %.3e  gx max %.3e (frame %d, ref32 there %.3e) median %.3e | 32768 frames fwd %.3f ms  fwd+dx %.3f ms"
          % (mode, seg, float(ey.max()), float(eg.max()), w, float(e32[w]), float(eg.median()), tf, tv))
