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
pp = model.preprocessing_layer
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gfeat = torch.randn(32768, spec.feature_dim(), device="cuda")
xbg = xb.clone().requires_grad_(True)
spec5 = S.get_spec("C5")
model5, _ = S.build_model(spec5)
model5 = model5.cuda()
x5 = S.make_frames(spec5, 8192, device="cuda", seed=3)
cot5 = torch.zeros(8192, 2, device="cuda"); cot5[:, 0] = 1
for staged, sbs, ctas in ((1, 0, 0), (2, 2, 0), (2, 3, 0)):
    os.environ["MOLANN_B200_STAGED"] = str(staged)
    os.environ["MOLANN_B200_SB_STAGES"] = str(sbs)
    os.environ["MOLANN_B200_SB_CTAS"] = str(ctas)
    xd = x.cuda().requires_grad_(True)
    y = model(xd)
    (gx,) = torch.autograd.grad(y, xd, cot.cuda())
    (gx2,) = torch.autograd.grad(model(xd), xd, cot.cuda())
    ey = frame_rel_err(y.detach().cpu(), y64); eg = frame_rel_err(gx.cpu(), gx64)
    with torch.no_grad():
        tpf = t(lambda: pp(xb))
        tf = t(lambda: model(xb))
        tf5 = t(lambda: model5(x5))
    feat = pp(xbg)
    tpb = t(lambda: torch.autograd.grad(feat, xbg, gfeat, retain_graph=True))
    tv = t(lambda: model.value_and_grad(xb, cotb))
    tv5 = t(lambda: model5.value_and_grad(x5, cot5))
    print("staged %d sb_stages %d ctas %d: y max %.3e gx max %.3e rerun-bitwise %s | C3 32768 frames: preprocess fwd %.3f ms bwd %.3f ms | model fwd %.3f ms fwd+dx %.3f ms | C5 8192 frames fwd %.3f fwd+dx %.3f ms"
          % (staged, sbs, ctas, float(ey.max()), float(eg.max()), bool(torch.equal(gx, gx2)), tpf, tpb, tf, tv, tf5, tv5))
