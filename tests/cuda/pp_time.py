"""Dev probe: stand-alone PreprocessingANN forward / backward rate on C2-sized frames (GPU)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from molann_b200 import synthetic as S
spec = S.get_spec("C2"); model, _ = S.build_model(spec); model = model.cuda(); pp = model.get_preprocessing_layer()
L = 1 << 20
x = S.make_frames(spec, L, device="cuda"); xg = x.clone().requires_grad_(True)
g = torch.randn(L, 30, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for tile in ("1", "0"):
    os.environ["MOLANN_B200_TILE"] = tile
    with torch.no_grad(): tf = t(lambda: pp(x))
    f = pp(xg); tb = t(lambda: torch.autograd.grad(f, xg, g, retain_graph=True))
    al = pp.align_layer
    with torch.no_grad(): ta = t(lambda: al(x))
    z = al(xg); go = torch.randn_like(z); tab = t(lambda: torch.autograd.grad(z, xg, go, retain_graph=True))
    print("TILE=%s AlignmentLayer   C2 2^20 frames: forward %.3f ms (%.2f G frames/s), backward %.3f ms (%.2f G frames/s)"
          % (tile, ta, L / ta / 1e6, tab, L / tab / 1e6))
    print("TILE=%s PreprocessingANN C2 2^20 frames: forward %.3f ms (%.2f G frames/s), backward %.3f ms (%.2f G frames/s)"
          % (tile, tf, L / tf / 1e6, tb, L / tb / 1e6))
