"""DEVELOPMENT TOOL: is a frame's result independent of its position in the batch?  (bitwise)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from molann_b200 import synthetic as S
spec = S.get_spec("C2")
spec.layer_dims = [30, 64, 2]
spec.activation = "relu"
spec.align_ix = None
model, _ = S.build_model(spec); model = model.cuda()
x = S.make_frames(spec, 512, device="cuda")
with torch.no_grad():
    y_full = model(x[:256].contiguous())
    single = torch.cat([model(x[j:j+1].contiguous()) for j in range(256)])
    print('full(256) vs singles: rows differing', int((y_full != single).any(dim=1).sum()))
    for a, n in ((1, 128), (1, 127), (2, 128), (1, 64), (3, 32), (1, 1), (1, 2), (1, 255)):
        ys = model(x[a:a+n].contiguous())
        print('slice', a, n, 'vs full rows differing', int((ys != y_full[a:a+n]).any(dim=1).sum()), 'vs singles', int((ys != single[a:a+n]).any(dim=1).sum()))
    # clone variants of the same slice
    s1 = x[1:129].contiguous(); s2 = x[1:129].clone(); s3 = torch.empty(128, 22, 3, device="cuda"); s3.copy_(x[1:129])
    print('data identical:', bool(torch.equal(s1, x[1:129])), 'ptr%16:', s1.data_ptr() % 16, s2.data_ptr() % 16, s3.data_ptr() % 16)
    print('contig vs clone vs copy equal:', bool(torch.equal(model(s1), model(s2))), bool(torch.equal(model(s1), model(s3))))
