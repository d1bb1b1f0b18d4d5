"""Timing probe of the fused wide forward kernel (C3 / C5) under its tunables.  python tests/cuda/wide_probe.py C3"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from molann_b200 import synthetic as S  # noqa: E402


def run(name, env):
    for k in list(os.environ):
        if k.startswith("MOLANN_B200_"):
            del os.environ[k]
    os.environ.update(env)
    spec = S.get_spec(name)
    L = spec.default_frames
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, L, device="cuda", seed=9)
    with torch.no_grad():
        for _ in range(3):
            y = model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            y = model(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fr = L / (ms * 1e-3)
    print("%s %-60s %.3f ms  %.1f M frames/s  frac %.3f  chk %.6f" % (
        name, env, ms, fr / 1e6, fr * spec.bytes_fwd() / 6558.1e9, float(y.double().sum())), flush=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["C3", "C5"]
    for name in names:
        run(name, {"MOLANN_B200_WIDE": "0"})
        run(name, {})
        ku2 = {"MOLANN_B200_WIDE_KU": "2"}
        for extra in ({}, {"MOLANN_B200_WIDE_STAGES": "4"}, {"MOLANN_B200_WIDE_STAGES": "3"},
                      {"MOLANN_B200_WIDE_STAGES": "3", "MOLANN_B200_WIDE_RING": "5"},
                      {"MOLANN_B200_WIDE_STAGES": "2", "MOLANN_B200_WIDE_RING": "6"},
                      {"MOLANN_B200_WIDE_STAGES": "4", "MOLANN_B200_WIDE_CDEPTH": "4"}):
            e = dict(ku2)
            e.update(extra)
            run(name, e)
