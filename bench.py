#!/usr/bin/env python
"""Benchmark of the molann hot path (BASELINE.json: frames/s, fwd and fwd+d/dx, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the fused align -> features -> MLP path over one device-resident batch of
synthetic frames (workload C2 = BASELINE.json configs[1]: alanine dipeptide, heavy-atom alignment +
30 position features -> MLP [30,64,64,2], 2**20 frames per GPU per step; weak scaling: every rank owns
its own frames, no data-path collective).  One JSON line is printed by rank 0:

  value       whole-job forward frames/s with inputs resident in HBM (CUDA events, max over ranks)
  fwd_dx      same for forward + gradient wrt coordinates (the biasing-force path)
  e2e         the same metric through the public API with HOST buffers (pinned H2D of every step's frames
              and D2H of its outputs inside the timed region, molann_b200.stream.HostPipeline)
  roofline    algorithmic bytes / kernel time vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline the reference's PyTorch CPU path on this box's host cores (bounded sample)

`--impl reference` times the reference's own CPU implementation (baseline/_ref = unmodified
zwpku/molann if installed, else the oracle port) on the same workload / metric.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames_per_sec_fwd"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-layers", action="store_true", help="skip the stand-alone AlignmentLayer / PreprocessingANN rates")
    ap.add_argument("--no-workloads", action="store_true",
                    help="default (C2) run: skip the C1 / C3 / C5 / C4 / latency entries of `workloads`")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            with open(path) as fh:
                d = json.load(fh)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, which):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(path):
        try:
            with open(path) as fh:
                return json.load(fh).get(workload, {}).get(which)
        except Exception:
            return None
    return None


class ClockSampler(object):
    """Samples SM clock + throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            from molann_b200.stream import nvml_handle
            self.h = nvml_handle(pynvml, index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        getter = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    if getter is not None:
                        mask = int(getter(self.h))
                        for bit, nm in names.items():
                            if mask & bit:
                                self.reasons.add(nm)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def close(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def reference_api():
    """-> (namespace, kind): the unmodified reference if importable, else the oracle port."""
    from types import SimpleNamespace
    from oracle.ref_loader import load_reference
    ref = load_reference()
    if ref is not None:
        ann, feature, root = ref
        return SimpleNamespace(Feature=feature.Feature, FeatureLayer=ann.FeatureLayer,
                               AlignmentLayer=ann.AlignmentLayer, PreprocessingANN=ann.PreprocessingANN,
                               MolANN=ann.MolANN, create_sequential_nn=ann.create_sequential_nn), "reference", root
    return None, "port", "oracle/restatement.py"


def cpu_model(spec):
    """CPU callable frames -> outputs for `spec` (reference modules, or the oracle restatement)."""
    import warnings
    from molann_b200 import synthetic as S
    api, kind, where = reference_api()
    if api is not None:
        model, _ = S.build_model(spec, api)

        def fn(x):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return model(x)
        return fn, kind, where
    from oracle import restatement as R
    import numpy as np
    ours, _ = S.build_model(spec)
    sd = ours.state_dict()
    nl = len(spec.layer_dims) - 1
    ws = [sd["ann_layers.%dth_layer.weight" % (k + 1)] for k in range(nl)]
    bs = [sd["ann_layers.%dth_layer.bias" % (k + 1)] for k in range(nl)]
    inp = list(spec.input_ix)
    tid = {"angle": 0, "bond": 1, "dihedral": 2, "position": 3}
    feats = [(tid[t], [inp.index(i) for i in ix]) for _, t, ix in spec.features]
    aidx = [inp.index(i) for i in spec.align_ix] if spec.align_ix is not None else None
    ref = torch.from_numpy(spec.positions[np.asarray(spec.align_ix)].copy()) if aidx is not None else torch.zeros(0, 3)
    ref = ref - ref.mean(0) if aidx is not None else ref

    def fn(x):
        return R.molann_forward(x, aidx, ref, feats, spec.use_angle_value, ws, bs, spec.activation)
    return fn, kind, where


def time_cpu(spec, frames_fwd, frames_dx, reps=3):
    """Best-of-`reps` CPU frames/s, forward (no_grad) and forward + d/dx (autograd)."""
    from molann_b200 import synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind, where = cpu_model(spec)
    x = S.make_frames(spec, max(frames_fwd, frames_dx), seed=spec.seed + 1)
    with torch.no_grad():
        fn(x[:1024])
    best_f = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        with torch.no_grad():
            fn(x[:frames_fwd])
        best_f = max(best_f, frames_fwd / (time.perf_counter() - t0))
    best_d = 0.0
    for _ in range(reps):
        xx = x[:frames_dx].clone().requires_grad_(True)
        t0 = time.perf_counter()
        y = fn(xx)
        torch.autograd.grad(y.sum(), xx)
        best_d = max(best_d, frames_dx / (time.perf_counter() - t0))
    return best_f, best_d, cores, kind, where


def cpu_sample_sizes(spec):
    n = spec.n_inp
    if n <= 64:
        return 1 << 19, 1 << 18
    if n <= 2000:
        return 4096, 512
    return 1024, 128


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU implementation on the same workload / metric (rank 0 only)."""
    if rank != 0:
        return
    from molann_b200 import synthetic as S
    spec = S.get_spec(args.workload)
    ff, fd = cpu_sample_sizes(spec)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind, where = cpu_model(spec)
    x = S.make_frames(spec, max(ff, fd), seed=spec.seed + 1)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 3))):
            fn(x[:ff])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(x[:ff])
        dt = time.perf_counter() - t0
    value = args.steps * ff / dt
    t0 = time.perf_counter()
    nd = max(1, args.steps // 4)
    for _ in range(nd):
        xx = x[:fd].clone().requires_grad_(True)
        torch.autograd.grad(fn(xx).sum(), xx)
    vdx = nd * fd / (time.perf_counter() - t0)
    sample = "%d frames/step fwd (no_grad), %d frames/step fwd+dx (autograd), %s, torch %d threads" % (
        ff, fd, where, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(spec, args.frames or spec.default_frames, args.gpus),
        "reference_sample_frames_per_step": ff,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "fwd_dx": {"value": vdx, "unit": UNIT},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# C4: data-parallel autoencoder training (BASELINE.json configs[3]): encoder = the C2 MolANN, decoder =
# create_sequential_nn([2,64,64,30]), loss = mean((dec(enc(x)) - preprocessing(x))^2), plain SGD, FIXED global batch of
# 2^20 frames split over the ranks (strong scaling), one flat allreduce of the MLP gradients per step.
# ------------------------------------------------------------------------------------------------------------
C4_GLOBAL = 1 << 20
C4_NOTE = ("C4: autoencoder training, encoder = C2 MolANN [30,64,64,2], decoder [2,64,64,30], MSE against the "
           "preprocessing output, SGD lr 1e-3, global batch 2^20 frames")


def c4_models(api=None):
    from molann_b200 import synthetic as S
    spec = S.get_spec("C2")
    enc, _ = S.build_model(spec, api)
    torch.manual_seed(404)
    create = (api or S.default_api()).create_sequential_nn
    dec = create([spec.out_dim(), 64, 64, spec.feature_dim()])
    return spec, enc, dec


def c4_cpu_step_rate(frames, steps, warmup=1):
    """The reference's modules (baseline/_ref, else the drop-in classes are NOT used: oracle port) training on CPU."""
    import warnings
    from molann_b200 import synthetic as S
    from molann_b200.train import AutoencoderStep
    api, kind, where = reference_api()
    if api is None:
        return None, kind, where
    spec, enc, dec = c4_models(api)
    step = AutoencoderStep(enc, dec, lr=1e-3)
    x = S.make_frames(spec, frames, seed=404)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(warmup):
            step.step(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            step.step(x)
        dt = time.perf_counter() - t0
    return steps * frames / dt, kind, where


def run_training(args, rank, local_rank, world):
    import torch.distributed as dist
    from molann_b200 import _lib
    from molann_b200 import synthetic as S
    from molann_b200.shard import frame_range
    from molann_b200.train import AutoencoderStep
    spec, enc, dec = c4_models()
    enc, dec = enc.cuda(), dec.cuda()
    n_global = args.frames * world if args.frames else C4_GLOBAL
    lo, hi = frame_range(n_global, rank, world)
    x = S.make_frames(spec, hi - lo, device="cuda", seed=404 + rank)
    trainer = AutoencoderStep(enc, dec, lr=1e-3, global_frames=n_global)
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    for _ in range(W):
        trainer.step(x)
    barrier()
    # eager: every kernel launched from the host
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for _ in range(K):
        loss = trainer.step(x)
    e1.record()
    barrier()
    sampler.pause()
    ms_eager = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0
    # the same step captured once as a CUDA graph (forward, backward, allreduce, SGD) and replayed
    graphed = bool(int(os.environ.get("MOLANN_BENCH_GRAPH", "1"))) and trainer.capture(x)
    if world > 1:
        flag = torch.tensor([1.0 if graphed else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        graphed = bool(flag.item() > 0.5)
    ms = ms_eager
    if graphed:
        for _ in range(W):
            trainer.replay()
        barrier()
        sampler.start()
        e0.record()
        for _ in range(K):
            loss = trainer.replay()
        e1.record()
        barrier()
        sampler.pause()
        ms = max_over_ranks(e0.elapsed_time(e1))
    # the collective alone (flat gradient buffer of this model), same stream, device-timed
    nparam = sum(p.numel() for p in trainer.params)
    flat = torch.zeros(nparam + 1, device="cuda")
    ar_ms = 0.0
    if world > 1:
        for _ in range(3):
            dist.all_reduce(flat)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            dist.all_reduce(flat)
        a1.record()
        barrier()
        ar_ms = max_over_ranks(a0.elapsed_time(a1)) / 20
    # end to end: this rank's shard arrives from pinned host memory every step, the loss is read back
    xh = x.cpu().pin_memory()
    ke = max(3, min(K, 10))
    run = trainer.replay if graphed else (lambda: trainer.step(x))
    if trainer._fused_args and int(os.environ.get("MOLANN_BENCH_TRAIN_PIPE", "1")):
        # the shard streams in pieces while the fused kernel works on the previous piece (AutoencoderStep.step_from_host)
        def e2e_step():
            return float(trainer.step_from_host(xh, chunks=int(os.environ.get("MOLANN_BENCH_TRAIN_CHUNKS", "16"))))
        e2e_api = "molann_b200.train.AutoencoderStep.step_from_host(pinned shard); float(loss)"
    else:
        def e2e_step():
            x.copy_(xh, non_blocking=True)
            return float(run())
        e2e_api = "x.copy_(pinned shard); molann_b200.train.AutoencoderStep.step(x); float(loss)"
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        last = e2e_step()
    torch.cuda.synchronize()
    ms_e = max_over_ranks(1e3 * (time.perf_counter() - t0))
    clocks = sampler.summary()
    sampler.close()
    h2d_bytes = int(xh.numel() * 4)
    del xh
    if rank != 0:
        return None
    peak, peak_src = peaks()
    value = n_global * K / (ms * 1e-3)
    bytes_per_frame = 12 * spec.n_inp                      # B_train = 12 n_inp per frame
    achieved = value * bytes_per_frame / world / 1e9
    line = {
        "metric": "frames_per_sec_train", "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": C4_NOTE, "n_inp": spec.n_inp, "global_batch_frames": n_global,
                   "frames_per_gpu_per_step": hi - lo, "trainable_parameters": nparam,
                   "l2_policy": "inputs larger than L2 (%.0f MB per step per GPU)" % ((hi - lo) * 12 * spec.n_inp / 1e6),
                   "parallelism": "data-parallel x%d, one flat sum-allreduce of %d floats per step (NCCL)"
                                  % (world, nparam + 1)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (ncu_traffic("C4", "train") if (hi - lo) == C4_GLOBAL and trainer._fused_args else None),
                     "peak_source": peak_src,
                     "note": "per GPU, algorithmic 12*n_inp bytes per frame; the step is FMA / shared-memory bound "
                             "(74 kFLOP per frame incl. weight gradients: %.1f TFLOP/s fp32 per GPU), see DESIGN.md 3.8"
                             % (value * 74e3 / world / 1e12)},
        "fp32": {"achieved_tflops": value * 74e3 / world / 1e12, "peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12,
                 "frac": (value * 74e3 / world / 1e12) / (148 * 128 * 2 * 1.965e9 / 1e12),
                 "note": "the step is bound by FP32 FMA work, not HBM: 74 kFLOP per frame against 148 SMs x 128 lanes x 2 "
                         "x 1.965 GHz; profiles/r5_train_ab.txt, DESIGN.md 3.8"},
        "train_path": ("fused_train_kernel (one kernel per step + plane reduction + SGD launch)"
                       if trainer._fused_args else "composed (fused encoder kernels + library decoder)"),
        "collective": ("none (1 GPU)" if world == 1 else
                       "one-shot allreduce over NVLink peer memory fused with SGD (train_allreduce_sgd_kernel)"
                       if getattr(trainer, "_peer", None) is not None else "NCCL allreduce of the flat gradient"),
        "gpu_launches": int(launches), "allreduce_ms": ar_ms, "final_loss": float(loss), "clocks": clocks,
        "cuda_graph": {"used": graphed, "eager_ms_per_step": ms_eager / K,
                       "note": "value is the graph replay when used (one launch per step replays gpu_launches / steps "
                               "of our kernels + the decoder's); gpu_launches counts the eager leg",
                       "capture_error": getattr(trainer, "_capture_error", None)},
        "e2e": {"value": n_global * ke / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "steps": ke, "ms_per_step": ms_e / ke, "last_loss": last,
                "api": e2e_api},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, kind, where = c4_cpu_step_rate(1 << 16, 3)
        if rate is not None:
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": "3 SGD steps on 65536 frames; %s; torch.set_num_threads(%d)"
                                              % (where, os.cpu_count() or 1)}
    return line


def workload_config(spec, frames, n_gpus):
    return {"workload": "%s: %s" % (spec.name, spec.note), "n_inp": spec.n_inp, "d_feat": spec.feature_dim(),
            "mlp": spec.layer_dims, "frames_per_gpu_per_step": frames,
            "bytes_per_frame_fwd": spec.bytes_fwd(), "bytes_per_frame_fwd_dx": spec.bytes_fwd_dx(),
            "l2_policy": "inputs larger than L2 (%.0f MB read per step per GPU, no flush needed)"
                         % (frames * 12 * spec.n_inp / 1e6),
            "parallelism": "frame-sharded x%d, no data-path collective" % n_gpus}


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner under
    NCCL_DEBUG=VERSION, torchrun notices) is sent to stderr; emit() writes to the original descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ------------------------------------------------------------------------------------------------------------
# inference workloads (C1, C2, C3, C5): forward and forward + d/dx, device-resident and end to end
# ------------------------------------------------------------------------------------------------------------
MIN_TIMED_MS = float(os.environ.get("MOLANN_BENCH_MIN_MS", "100"))   # every leg is warmed up for and timed over at
                                                                      # least this long (clock ramp, noise); 0 for ncu runs
MAX_REPS = 40


class Dist(object):
    """barrier / reductions over the ranks of this job (no-ops for one rank)."""

    def __init__(self, world):
        self.world = world
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def gather(self, value):
        """-> list of `value` over the ranks (every rank gets it)."""
        if self.world == 1:
            return [float(value)]
        t = torch.zeros(self.world, device="cuda", dtype=torch.float64)
        t[self.dist.get_rank()] = float(value)
        self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]


def timed_leg(step, K, W, dd, sampler=None):
    """Warm up (>= W steps and >= MIN_TIMED_MS), then time repetitions of EXACTLY K steps -- each bracketed by a
    barrier + synchronize on both sides, CUDA events on the launching stream -- until MIN_TIMED_MS of device time
    has been measured.  The leg's figure is the MEDIAN repetition of the max-over-ranks time.  Returns
    (ms for K steps, launches per K steps, stats)."""
    from molann_b200 import _lib
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(W):
        step()
    dd.barrier()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    dd.barrier()
    guess = max(e0.elapsed_time(e1), 1e-3)
    extra = int(min(50.0, MIN_TIMED_MS / guess))            # time-based part of the warm-up
    if dd.world > 1:
        extra = int(max(dd.gather(extra)))
    for _ in range(extra * K):
        step()
    reps, total, launches, per_rank = [], 0.0, 0, []
    while True:
        dd.barrier()
        l0 = _lib.launch_count()
        if sampler is not None:
            sampler.start()
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        dd.barrier()
        if sampler is not None:
            sampler.pause()
        launches = _lib.launch_count() - l0
        ranks = dd.gather(e0.elapsed_time(e1))
        per_rank.append(ranks)
        reps.append(max(ranks))
        total += max(ranks)
        if total >= MIN_TIMED_MS or len(reps) >= MAX_REPS:
            break
    order = sorted(range(len(reps)), key=lambda i: reps[i])
    mid = order[len(order) // 2]
    pr = per_rank[mid]
    stats = {"repetitions": len(reps), "ms_min": min(reps), "ms_median": reps[mid], "ms_max": max(reps),
             "per_rank_ms_of_median_rep": {"min": min(pr), "median": sorted(pr)[len(pr) // 2], "max": max(pr)},
             "rule": "median over repetitions of max over ranks; each repetition = K steps between barriers"}
    return reps[mid], launches, stats


def measure_workload(name, args, rank, local_rank, world, dd, full):
    """One inference workload on this job's ranks.  `full`: also the stand-alone layers, both end-to-end legs and
    the CPU baseline (the headline workload); otherwise forward, forward + d/dx and the forward end-to-end leg."""
    from molann_b200 import synthetic as S
    from molann_b200.stream import HostPipeline
    spec = S.get_spec(name)
    frames = (args.frames if (args.frames and full) else 0) or spec.default_frames
    model, _ = S.build_model(spec)
    model = model.cuda()
    x = S.make_frames(spec, frames, device="cuda", seed=spec.seed + rank)
    cot = torch.zeros(frames, spec.out_dim(), device="cuda")
    cot[:, 0] = 1.0                                        # d y_0 / dx: one collective-variable force
    K, W = args.steps, max(args.warmup, 3)

    def fwd_step():
        with torch.no_grad():
            return model(x)

    def fwd_dx_step():
        # outputs + coordinate gradient in ONE pass (MolANN.value_and_grad, the biasing-force entry)
        return model.value_and_grad(x, cot)

    sampler = ClockSampler(local_rank)
    ms_f, launches_f, st_f = timed_leg(fwd_step, K, W, dd, sampler)
    ms_d, launches_d, st_d = timed_leg(fwd_dx_step, K, W, dd, sampler)
    clocks = sampler.summary()
    sampler.close()
    # cross-rank result check: every rank ALSO evaluates rank 0's first frames; all ranks must agree bit for bit
    # (frames are independent, so a shard's output may not depend on the GPU it ran on)
    n_chk = min(frames, 4096)
    x0 = S.make_frames(spec, n_chk, device="cuda", seed=spec.seed)
    with torch.no_grad():
        y0 = model(x0)
    chk = float(y0.double().sum())
    chks = dd.gather(chk)
    checksum = {"frames": n_chk, "value": chk, "identical_across_ranks": bool(all(c == chks[0] for c in chks))}
    del x0, y0

    via_autograd = None
    layers = None
    if full:
        xg = x.clone().requires_grad_(True)

        def fwd_dx_autograd_step():
            y = model(xg)
            (g,) = torch.autograd.grad(y, xg, cot)
            return g

        ms_da, launches_da, _ = timed_leg(fwd_dx_autograd_step, K, W, dd)
        via_autograd = {"value": world * frames * K / (ms_da * 1e-3), "ms_per_step": ms_da / K,
                        "gpu_launches": int(launches_da), "api": "y = model(x); torch.autograd.grad(y, x, cotangent)"}
        # ---- the stand-alone layers of the same model ----
        pp = model.get_preprocessing_layer()
        if not args.no_layers and getattr(pp, "align_layer", None) is not None and hasattr(pp.align_layer, "ref_x"):
            n3b, db = 12 * spec.n_inp, 4 * spec.feature_dim()
            gfe = torch.zeros(frames, spec.feature_dim(), device="cuda")
            gfe[:, 0] = 1.0

            def rate(step, nbytes):
                ms_l, _, _ = timed_leg(step, K, W, dd)
                v = world * frames * K / (ms_l * 1e-3)
                return {"value": v, "unit": UNIT, "ms_per_step": ms_l / K, "bytes_per_frame": nbytes,
                        "frac_of_hbm_peak": v / world * nbytes / 1e9 / peaks()[0]}

            def al_f():
                with torch.no_grad():
                    return pp.align_layer(x)

            def pp_f():
                with torch.no_grad():
                    return pp(x)

            za = pp.align_layer(xg)
            fa = pp(xg)
            layers = {
                "AlignmentLayer.forward": rate(al_f, 2 * n3b),
                "AlignmentLayer.backward": rate(lambda: torch.autograd.grad(za, xg, x, retain_graph=True), 3 * n3b),
                "PreprocessingANN.forward": rate(pp_f, n3b + db),
                "PreprocessingANN.backward": rate(lambda: torch.autograd.grad(fa, xg, gfe, retain_graph=True),
                                                  2 * n3b + db),
            }
            del za, fa, gfe
        del xg

    # ---- end to end through the public API with host buffers ----
    e2e = e2e_dx = None
    if not args.no_e2e:
        xh = x.cpu().pin_memory()
        yh = torch.empty(frames, spec.out_dim()).pin_memory()
        chunk = int(os.environ.get("MOLANN_BENCH_E2E_CHUNK", max(1 << 10, frames // 8)))
        pipe = HostPipeline(model, spec.n_inp, spec.out_dim(), chunk_frames=chunk)
        ke = max(3, min(K, 10))
        for _ in range(2):
            pipe.run(xh, yh)
        dd.barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            pipe.run(xh, yh)                              # returns once the results are in host memory
        torch.cuda.synchronize()
        ranks = dd.gather(1e3 * (time.perf_counter() - t0))
        ms_e = max(ranks)
        e2e = {"value": world * frames * ke / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "steps": ke, "ms_per_step": ms_e / ke,
               "api": "molann_b200.stream.HostPipeline(MolANN).run(pinned x, pinned y)",
               "result_checksum": float(yh[:: max(1, frames // 1024)].double().sum()),
               "per_rank_h2d_GBps": [pipe.h2d_bytes * ke / (r * 1e-3) / 1e9 for r in ranks]}
        if full:
            # the same leg over the int16 wire format (half the H2D bytes; decode = one small kernel per chunk)
            from molann_b200.stream import quantize_frames
            qh, origin, res = quantize_frames(xh)
            qh = qh.pin_memory()
            for _ in range(2):
                pipe.run_wire(qh, origin, res, yh)
            dd.barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                pipe.run_wire(qh, origin, res, yh)
            torch.cuda.synchronize()
            ms_w = max(dd.gather(1e3 * (time.perf_counter() - t0)))
            e2e["int16_wire"] = {"value": world * frames * ke / (ms_w * 1e-3), "unit": UNIT,
                                 "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                                 "ms_per_step": ms_w / ke, "resolution": res,
                                 "api": "HostPipeline.run_wire(quantize_frames(x)): lossy transport, NOT the parity "
                                        "default (outputs are exact functions of the decoded coordinates)"}
            del qh
            coth = cot.cpu().pin_memory()
            gxh = torch.empty(frames, spec.n_inp, 3).pin_memory()
            pipe.run(xh, yh, coth, gxh)
            dd.barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                pipe.run(xh, yh, coth, gxh)
            torch.cuda.synchronize()
            ms_ed = max(dd.gather(1e3 * (time.perf_counter() - t0)))
            e2e_dx = {"value": world * frames * ke / (ms_ed * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "steps": ke,
                      "ms_per_step": ms_ed / ke}
            del coth, gxh
        del xh, yh, pipe

    out = None
    if rank == 0:
        peak, peak_src = peaks()

        def roofline(bytes_per_frame, ms_total, launches, which):
            # SURVEY 8(d) honesty guard: credit min(algorithmic bytes, DRAM bytes ncu measured for this launch
            # shape), so a kernel is never credited for bytes it did not move.  The ncu figure only applies to the
            # launch shape it was captured on.
            per_launch_s = ms_total * 1e-3 / K
            alg = bytes_per_frame * frames
            traffic = ncu_traffic(spec.name, which) if frames == spec.default_frames else None
            credited = min(alg, traffic) if traffic else alg
            achieved = credited / per_launch_s / 1e9
            return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                    "credited_bytes_per_launch": credited, "kernel_launches_per_step": launches / K,
                    "frac_of_nominal_8TBs": achieved / 8000.0}

        out = {
            "metric": METRIC, "value": world * frames * K / (ms_f * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_f / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(spec, frames, world),
            "roofline": roofline(spec.bytes_fwd(), ms_f, launches_f, "fwd"),
            "gpu_launches": int(launches_f), "clocks": clocks, "timing": st_f,
            "fwd_dx": {"value": world * frames * K / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d / K,
                       "gpu_launches": int(launches_d), "api": "MolANN.value_and_grad(x, cotangent)",
                       "roofline": roofline(spec.bytes_fwd_dx(), ms_d, launches_d, "fwd_dx"), "e2e": e2e_dx,
                       "timing": st_d, "via_autograd": via_autograd},
            "e2e": e2e, "cross_rank_checksum": checksum,
        }
        if layers is not None:
            out["layers"] = layers
        if world == 1 and not args.no_cpu_baseline:
            ff, fd = cpu_sample_sizes(spec)
            cf, cd, cores, kind, where = time_cpu(spec, ff, fd, reps=3 if full else 2)
            out["cpu_baseline"] = {"value": cf, "unit": UNIT, "cores": cores, "kind": kind,
                                   "sample": "best of %d: %d frames fwd (no_grad), %d frames fwd+dx (autograd); %s; "
                                             "torch.set_num_threads(%d)" % (3 if full else 2, ff, fd, where, cores),
                                   "fwd_dx_value": cd}
    del model, x, cot
    torch.cuda.empty_cache()
    return out


def latency_probe(dd):
    """What an MD plugin sends: L = 1 ... 128 frames per call.  Microseconds per call (host wall clock around a
    synchronised loop of calls through the module API), forward and value_and_grad, workload C2."""
    from molann_b200 import synthetic as S
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    model = model.cuda()
    res = {}
    for L in (1, 32, 128):
        x = S.make_frames(spec, L, device="cuda", seed=5)
        cot = torch.ones(L, spec.out_dim(), device="cuda")
        with torch.no_grad():
            for _ in range(20):
                model(x)
                model.value_and_grad(x, cot)
            torch.cuda.synchronize()
            n = 200
            t0 = time.perf_counter()
            for _ in range(n):
                model(x)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(n):
                model.value_and_grad(x, cot)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            # one call at a time, result needed on the host before the next MD step can start
            t3 = time.perf_counter()
            for _ in range(n):
                y, g = model.value_and_grad(x, cot)
                y.cpu()
            t4 = time.perf_counter()
            t5 = time.perf_counter()
            for _ in range(n):
                y, jac = model.value_and_jacobian(x)
                y.cpu()
            t6 = time.perf_counter()
        res["L=%d" % L] = {"forward_us_per_call": 1e6 * (t1 - t0) / n, "value_and_grad_us_per_call": 1e6 * (t2 - t1) / n,
                           "value_and_grad_sync_us_per_call": 1e6 * (t4 - t3) / n,
                           "value_and_jacobian_sync_us_per_call": 1e6 * (t6 - t5) / n}
    # the reference's way on the host cores: y = model(x), then one autograd call per output
    try:
        import warnings
        api, kind, where = reference_api()
        if api is not None:
            cpu_model_, _ = S.build_model(spec, api)
            for L in (1, 32, 128):
                xc = S.make_frames(spec, L, seed=5).requires_grad_(True)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    for _ in range(3):
                        yc = cpu_model_(xc)
                        [torch.autograd.grad(yc[:, o].sum(), xc, retain_graph=True) for o in range(yc.shape[1])]
                    t0 = time.perf_counter()
                    nrep = 50
                    for _ in range(nrep):
                        yc = cpu_model_(xc)
                        [torch.autograd.grad(yc[:, o].sum(), xc, retain_graph=True) for o in range(yc.shape[1])]
                    res["L=%d" % L]["reference_cpu_value_and_jacobian_us_per_call"] = 1e6 * (time.perf_counter() - t0) / nrep
            res["reference"] = "%s (%s), %d torch threads" % (kind, where, torch.get_num_threads())
    except Exception as exc:  # noqa: BLE001
        res["reference_error"] = repr(exc)
    # throughput of the all-outputs Jacobian at the bench size
    xb = S.make_frames(spec, 1 << 20, device="cuda", seed=6)
    for _ in range(3):
        model.value_and_jacobian(xb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        model.value_and_jacobian(xb)
    e1.record()
    torch.cuda.synchronize()
    res["value_and_jacobian_2^20_frames"] = {"ms_per_call": e0.elapsed_time(e1) / 10,
                                            "frames_per_s": (1 << 20) * 10 / (e0.elapsed_time(e1) * 1e-3),
                                            "planes": spec.out_dim(), "launches_per_call": 1}
    del xb
    res["note"] = ("C2 model, back-to-back calls through MolANN.forward / MolANN.value_and_grad (host wall clock / "
                   "calls); *_sync: each call followed by a device->host read of y")
    return res


def compact(line):
    """What the default line keeps of a secondary workload."""
    if line is None:
        return None
    keep = {"value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"],
            "frames_per_gpu_per_step": line["config"]["frames_per_gpu_per_step"],
            "gpu_launches_per_step": line["gpu_launches"] / float(line["steps"]),
            "roofline": {k: line["roofline"][k] for k in ("frac", "achieved", "traffic", "algorithmic_bytes_per_launch")},
            "fwd_dx": {"value": line["fwd_dx"]["value"], "ms_per_step": line["fwd_dx"]["ms_per_step"],
                       "gpu_launches_per_step": line["fwd_dx"]["gpu_launches"] / float(line["steps"]),
                       "roofline": {k: line["fwd_dx"]["roofline"][k]
                                    for k in ("frac", "achieved", "traffic", "algorithmic_bytes_per_launch")}},
            "e2e": line.get("e2e"), "cross_rank_checksum": line.get("cross_rank_checksum"),
            "workload": line["config"]["workload"]}
    if "cpu_baseline" in line:
        keep["cpu_baseline"] = line["cpu_baseline"]
    return keep


def main():
    args = parse_args()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.workload == "C4":
            if rank == 0:
                torch.set_num_threads(os.cpu_count() or 1)
                rate, kind, where = c4_cpu_step_rate(1 << 16, max(1, args.steps), 1)
                emit({"impl": "reference", "metric": "frames_per_sec_train", "value": rate, "unit": UNIT,
                                  "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                                  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                                  "data": "synthetic", "config": {"workload": C4_NOTE},
                                  "cpu_baseline": {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1,
                                                   "kind": kind, "sample": "SGD steps on 65536 frames; " + where},
                                  "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0}})
            return
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- molann_b200 has no CPU path (use --impl reference for the CPU arm)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        # host buffers of the end-to-end leg live on the NUMA node next to this rank's GPU
        from molann_b200.stream import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dd = Dist(world)
    torch.set_num_threads(os.cpu_count() or 1)

    if args.workload == "C4":
        line = run_training(args, rank, local_rank, world)
        if rank == 0:
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    line = measure_workload(args.workload, args, rank, local_rank, world, dd, full=True)
    # the other BASELINE configs ride along in the default line (C2) so the driver's record covers every one of them
    others = {}
    if args.workload == "C2" and not args.frames and not args.no_workloads:
        for name in ("C1", "C3", "C5"):
            others[name] = compact(measure_workload(name, args, rank, local_rank, world, dd, full=False))
        c4 = run_training(args, rank, local_rank, world)
        if c4 is not None:
            others["C4"] = {"metric": c4["metric"], "value": c4["value"], "unit": c4["unit"], "scaling": c4["scaling"],
                            "ms_per_step": c4["ms_per_step"], "allreduce_ms": c4["allreduce_ms"],
                            "gpu_launches": c4["gpu_launches"], "cuda_graph": c4["cuda_graph"]["used"],
                            "e2e": c4["e2e"], "final_loss": c4["final_loss"], "workload": c4["config"]["workload"],
                            "cpu_baseline": c4.get("cpu_baseline"), "train_path": c4.get("train_path"),
                            "collective": c4.get("collective"), "roofline": c4.get("roofline"), "fp32": c4.get("fp32")}
        if world == 1:
            others["latency_C2"] = latency_probe(dd)
    if rank == 0:
        if others:
            line["workloads"] = others
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
