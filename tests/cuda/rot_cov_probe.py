import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from molann_b200 import synthetic as S
from helpers import oracle_model, oracle_value_and_grad, frame_rel_err
for name in ("C3", "C5"):
    spec = S.get_spec(name); model, _ = S.build_model(spec); model = model.cuda()
    Ls = 256
    x = S.make_frames(spec, Ls, device="cuda", seed=9)
    cot = torch.randn(Ls, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    y, g1 = model.value_and_grad(x, cot)
    Rm = S.random_rotations(Ls, torch.Generator(device="cuda").manual_seed(7), "cuda")
    x2 = (torch.bmm(x, Rm) + 5.0).contiguous()
    y2, g2 = model.value_and_grad(x2, cot)
    g1r = torch.bmm(g1, Rm)
    print(name, "y inv err", float((y2 - y).abs().max()), "scale", float(y.abs().max()), "| gx cov err", float((g2 - g1r).abs().max()), "scale", float(g1.abs().max()))
    sd = model.state_dict()
    ws = [sd["ann_layers.%dth_layer.weight" % k].cpu() for k in (1, 2, 3)]; bs = [sd["ann_layers.%dth_layer.bias" % k].cpu() for k in (1, 2, 3)]
    xs, cs = x[:32].cpu(), cot[:32].cpu()
    y64, gx64 = oracle_value_and_grad(oracle_model(spec, ws, bs), xs, cs)
    x2s = x2[:32].cpu()
    y64b, gx64b = oracle_value_and_grad(oracle_model(spec, ws, bs), x2s, cs)
    y32, gx32 = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), xs, cs, torch.float32)
    y32b, gx32b = oracle_value_and_grad(oracle_model(spec, ws, bs, torch.float32), x2s, cs, torch.float32)
    print("   ours vs fp64: x", float(frame_rel_err(g1[:32].cpu(), gx64).max()), "x2", float(frame_rel_err(g2[:32].cpu(), gx64b).max()),
          "| ref32 vs fp64: x", float(frame_rel_err(gx32, gx64).max()), "x2", float(frame_rel_err(gx32b, gx64b).max()),
          "| ref32 covariance err", float((gx32b - torch.bmm(gx32, Rm[:32].cpu())).abs().max() / gx32.abs().max()))
