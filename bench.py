#!/usr/bin/env python
"""bench.py — MLA train samples/sec (CREMA-D AV, ResNet-18, --gs_flag --dynamic), BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework
    python bench.py --impl reference ...                           # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A step = one alternating training step (audio turn + visual turn: forward of both encoders,
per turn head fwd+bwd, encoder backward, GS projection, SGD) over one synthetic batch of 64
samples per GPU (spectrogram 1x257x188, 2 frames 3x224x224, 6 classes).
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
FLOP_PER_SAMPLE_STEP = 32.47e9          # ResNet-18 pair fwd 10.823 GFLOP x3 (fwd+dgrad+wgrad)
WORKLOAD = "configs[1]: CREMA-D-shaped MLA (--gs_flag --dynamic) train step, batch 64 per GPU, ResNet-18 x2"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-sweep", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the informational transformer-path timings (N=1 only)")
    p.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager legs (N=1 only)")
    p.add_argument("--no-tf32-leg", action="store_true", help="skip the TF32-operand re-run of the step (N=1 only)")
    p.add_argument("--tf32-leg", action="store_true", help=argparse.SUPPRESS)     # internal: child process of the TF32 leg
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def profile_summary():
    """DRAM traffic per launch of the two reported kernels, from the committed ncu --set full summaries."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.3 * max(mx)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_args():
    return argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=True, lorb="base", modal3=False, clip=False)


# ------------------------------------------------------------------------------ reference arm
def run_reference(a, rank):
    """The reference's own CPU implementation of the path, as restated by the oracle port (the
    reference is Python: it cannot travel to the GPU box and may not be copied into the repo).
    All host threads torch can use; each step is a bounded sample (B_s <= 64 samples)."""
    if rank != 0:
        return
    import torch
    from oracle import mla_oracle as orc
    import mla_b200
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(make_args()).apply(mla_b200.weight_init)      # parameter container only (CPU init)
    o = orc.AVOracle(net.state_dict(), force_projection=True)
    # the full B = 64 batch of the metric's configuration whenever the whole run stays within ~4 minutes (2.4 s per step
    # on 16 cores: 25 steps = 60 s); otherwise a bounded sample of it
    spec, image, label = orc.synthetic_av_batch(4, 7)
    t0 = time.perf_counter(); o.train_step(spec, image, label, 0, 1); per_sample = (time.perf_counter() - t0) / 4
    budget = 240.0 / max(1, a.steps + a.warmup)
    bs = int(max(2, min(BATCH, budget / per_sample)))
    spec, image, label = orc.synthetic_av_batch(bs, 1)
    for i in range(a.warmup):
        o.train_step(spec, image, label, i, a.steps + a.warmup)
    t0 = time.perf_counter()
    for i in range(a.steps):
        o.train_step(spec, image, label, i, a.steps)
    dt = time.perf_counter() - t0
    val = bs * a.steps / dt
    out = {"impl": "reference", "metric": "MLA train samples/sec (CREMA-D AV, ResNet-18)", "value": val,
           "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "global_batch": bs, "sample_batch": bs, "device": "host CPU",
                      "same_batch_as_native_arm": bs == BATCH},
           "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": "%d timed steps of the oracle's alternating step on %d-sample batches" % (a.steps, bs)},
           "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_baseline():
    import torch
    from oracle import mla_oracle as orc
    import mla_b200
    torch.set_num_threads(os.cpu_count() or 1)
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(make_args()).apply(mla_b200.weight_init)
    o = orc.AVOracle(net.state_dict(), force_projection=True)
    bs = BATCH                                     # the metric's own batch: ~2.4 s per step on 16 host cores
    spec, image, label = orc.synthetic_av_batch(bs, 1)
    o.train_step(spec, image, label, 0, 3)
    t0 = time.perf_counter()
    n = 0
    while n < 2 or (time.perf_counter() - t0 < 10 and n < 4):
        o.train_step(spec, image, label, n + 1, 8)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": bs * n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of the oracle's alternating step on %d-sample batches (1 warm-up)" % (n, bs)}


# --------------------------------------------------------------------------------- native arm
SWEEP_D, SWEEP_B, SWEEP_C = (512, 768, 1024, 2048), (64, 256, 1024, 4096), (6, 101)     # BASELINE.json configs[4]


def _time_launch(torch, make_set, run, nbytes, launches=24, warm=3):
    """CUDA-event time of one launch of run(set) (trimmed mean over >= 24 launches). Cold inputs without evicting the kernel's code: the launches walk
    over `nsets` independent argument sets whose total footprint exceeds twice the 126 MB L2 (whenever nsets <= 128 allows
    it: the remaining points are latency-bound and labelled so), so every launch reads its inputs from HBM. Every timed
    launch is preceded by a ~60 us spin kernel, which lets the host run ahead: event, launch and event are queued before
    the GPU reaches them, and the events bracket the launch only (no host enqueue time, no flush traffic)."""
    nsets = max(2, min(128, -(-(300 << 20) // max(nbytes, 1)) + 1))
    sets = [make_set() for _ in range(nsets)]
    for i in range(min(warm, nsets)):
        run(sets[i])
    torch.cuda.synchronize()
    n = max(launches, nsets)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        torch.cuda._sleep(120000)
        ev[i][0].record()
        run(sets[i % nsets])
        ev[i][1].record()
    torch.cuda.synchronize()
    # CUDA-event timestamps tick every ~2 us on this GPU: the median of single-launch times snaps to that grid (36.9 / 38.9 us
    # for the same kernel from run to run), so the estimate is the mean of the central half of the samples instead
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    mid = ts[len(ts) // 4: len(ts) - len(ts) // 4]
    return sum(mid) / len(mid) * 1e-3, nsets


def _point(pk, t, nbytes, **kw):
    # a launch whose algorithmic bytes would take < 3 us at the HBM peak cannot be HBM-bound: launch latency, grid-wide
    # synchronisation and L2 residency decide its time; such points are labelled, not hidden
    t, nsets = t
    d = dict(kw, us=t * 1e6, bytes=nbytes, gbs=nbytes / t / 1e9, frac=nbytes / t / 1e9 / pk["hbm"], sets=nsets)
    d["regime"] = "latency" if nbytes / (pk["hbm"] * 1e9) < 3e-6 else "bandwidth"
    return d


def gs_sweep(torch, ops, pk):
    """GSPlugin HBM GB/s (second half of BASELINE.json's metric) over the full configs[4] grid: algorithmic bytes
    4*(B*D + 2*D*D + 2*C*D) / CUDA-event time of one launch, inputs cold (rotating argument sets, see _time_launch)."""
    from mla_b200.gs_plugin import GSPlugin
    dev = torch.device("cuda")
    alpha = GSPlugin.alpha(1, 10)
    pts = []
    for D in SWEEP_D:
        for B in SWEEP_B:
            for C in SWEEP_C:
                nbytes = 4 * (B * D + 2 * D * D + 2 * C * D)
                mk = lambda: (torch.eye(D, device=dev), torch.randn(C, D, device=dev), torch.randn(B, D, device=dev).relu())
                t = _time_launch(torch, mk, lambda s: ops.gs_project(s[0], s[1], alpha, feat=s[2]), nbytes)
                pts.append(_point(pk, t, nbytes, B=B, D=D, C=C))
    return pts


def gs_beyond_sweep(torch, ops, pk):
    """Informational, outside configs[4]: the same kernel at batch sizes where the feature stream dominates the fixed costs
    (cooperative launch, three grid barriers, the dependent P phases) that bound the sweep's 67 MB top point."""
    from mla_b200.gs_plugin import GSPlugin
    dev = torch.device("cuda")
    alpha = GSPlugin.alpha(1, 10)
    pts = []
    for B in (16384, 65536):
        D, C = 2048, 6
        nbytes = 4 * (B * D + 2 * D * D + 2 * C * D)
        mk = lambda: (torch.eye(D, device=dev), torch.randn(C, D, device=dev), torch.randn(B, D, device=dev).relu())
        t = _time_launch(torch, mk, lambda s: ops.gs_project(s[0], s[1], alpha, feat=s[2]), nbytes, launches=12)
        pts.append(_point(pk, t, nbytes, B=B, D=D, C=C))
    return pts


def producer_leg(torch, pk):
    """Informational: the visual dataset tuple producer — 64 samples x 2 decoded 360 x 480 RGB frames
    (uint8, already resident in HBM) -> [64, 3, 2, 224, 224] fp32, the reference's evaluation transform
    (dataset/dataset.py:133-138). Algorithmic bytes: frames read once + output written once."""
    import numpy as np
    from mla_b200 import ops
    dev = torch.device("cuda")
    B, T, H, W, S = 64, 2, 360, 480, 224
    nbytes = B * T * H * W * 3 + B * 3 * T * S * S * 4
    desc = np.zeros((B * T, 14), np.int32)
    for n in range(B * T):
        desc[n] = (n * H * W * 3, 0, H, W, 0, 0, H, W, 0, n, S, S, 0, 0)

    def mk():
        return (torch.randint(0, 256, (B * T * H * W * 3,), dtype=torch.uint8, device=dev), torch.from_numpy(desc).to(dev),
                torch.empty(B, 3, T, S, S, device=dev))
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    t, nsets = _time_launch(torch, mk, lambda s: ops.frames_to_batch(s[0], s[1], B, T, S, H, mean, std, out=s[2]), nbytes,
                            launches=12)
    return {"kernels": "frame_coeffs_kernel + frame_resample_kernel", "us": t * 1e6, "bytes": nbytes,
            "gbs": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / pk["hbm"], "samples_per_s": B / t, "sets": nsets,
            "workload": "64 x 2 frames 360x480x3 uint8 -> [64,3,2,224,224] fp32, bit-identical to torchvision on PIL"}


def head_sweep(torch, ops, pk):
    """Shared head forward + backward (main.py:432-435): algorithmic bytes 4*(2*B*D + 3*C*D + 2*B*C)."""
    dev = torch.device("cuda")
    pts = []
    for D in SWEEP_D:
        for B in SWEEP_B:
            for C in SWEEP_C:
                nbytes = 4 * (2 * B * D + 3 * C * D + 2 * B * C)
                mk = lambda: (torch.randn(B, D, device=dev).relu(), torch.randn(C, D, device=dev) * 0.05,
                              torch.zeros(C, device=dev), torch.randint(0, C, (B,), device=dev), {})
                t = _time_launch(torch, mk, lambda s: ops.head_ce(s[0], s[1], s[2], s[3], out=s[4]), nbytes)
                pt = _point(pk, t, nbytes, B=B, D=D, C=C)
                # three B x D x C products in fp32 on the CUDA cores (fp32 accuracy is part of the contract): at C = 101 the
                # head is bound by the FMA rate (148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s), not by HBM
                flops = 6.0 * B * D * C
                pt["fp32_tflops"] = flops / t[0] / 1e12
                if pt["regime"] == "bandwidth" and flops / 74.4e12 > nbytes / (pk["hbm"] * 1e9):
                    pt["regime"] = "fp32-compute"
                pts.append(pt)
    return pts


def fusion_sweep(torch, ops, pk):
    """Test-time entropy fusion + argmax + counters (main.py:65-106, 640-676): bytes 4*(M+1)*B*C + 8*B."""
    dev = torch.device("cuda")
    pts = []
    for M in (2, 3):
        for B in SWEEP_B:
            for C in SWEEP_C:
                nbytes = 4 * (M + 1) * B * C + 8 * B
                mk = lambda: ([torch.randn(B, C, device=dev) for _ in range(M)], torch.randint(0, C, (B,), device=dev),
                              torch.zeros(M + 1, C, dtype=torch.int64, device=dev), torch.zeros(C, dtype=torch.int64, device=dev))
                t = _time_launch(torch, mk, lambda s: ops.fuse_eval(s[0], s[1], hits=s[2], num=s[3]), nbytes)
                pts.append(_point(pk, t, nbytes, M=M, B=B, C=C))
    return pts


# ----------------------------------------------------------------------------- PyTorch-eager legs
def eager_baseline(torch, steps=5, warm=2):
    """The reference's step (main.py:127-164, 419-476) executed by PyTorch eager on this GPU — cuDNN / cuBLAS / ATen, the
    reference's per-turn optimizer.step() and per-step .item() reads — in three arithmetic modes: torch's defaults (cuDNN
    TF32: what the reference runs as published), strict fp32, and autocast-bf16 with channels_last tensors (the usual
    2-byte eager configuration; narrower than the reference, listed for context next to the 2-byte kernels)."""
    import torch.nn.functional as F
    import mla_b200
    dev = torch.device("cuda")
    res = {}
    layers = ((64, 1), (128, 2), (256, 2), (512, 2))

    def bn(x, sd, p):
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], True, 0.1,
                            1e-5)

    def resnet(sd, p, x, visual):                       # models/backbone.py:142-160, 36-52
        if visual:
            B, C, T, H, W = x.shape
            x = x.permute(0, 2, 1, 3, 4).contiguous().view(B * T, C, H, W)
        x = F.max_pool2d(F.relu(bn(F.conv2d(x, sd[p + "conv1.weight"], None, 2, 3), sd, p + "bn1")), 3, 2, 1)
        inpl = 64
        for li, (planes, stride) in enumerate(layers, start=1):
            for bi in range(2):
                st = stride if bi == 0 else 1
                q = "%slayer%d.%d." % (p, li, bi)
                idn = x
                o = F.relu(bn(F.conv2d(x, sd[q + "conv1.weight"], None, st, 1), sd, q + "bn1"))
                o = bn(F.conv2d(o, sd[q + "conv2.weight"], None, 1, 1), sd, q + "bn2")
                if bi == 0 and (st != 1 or inpl != planes):
                    idn = bn(F.conv2d(x, sd[q + "downsample.0.weight"], None, st), sd, q + "downsample.1")
                x = F.relu(o + idn)
                inpl = planes
        return x

    gen = torch.Generator().manual_seed(1)
    spec = torch.randn(BATCH, 257, 188, generator=gen).to(dev)
    image = torch.randn(BATCH, 3, 2, 224, 224, generator=gen).to(dev)
    label = torch.randint(0, 6, (BATCH,), generator=gen).to(dev)
    prev = torch.backends.cudnn.allow_tf32
    for mode in ("tf32", "fp32", "amp_bf16_channels_last"):
        try:
            torch.backends.cudnn.allow_tf32 = mode != "fp32"
            mla_b200.setup_seed(0)
            net = mla_b200.AVClassifier(make_args()).apply(mla_b200.weight_init)      # parameter container (same init)
            sd = {k: v.detach().to(dev) for k, v in net.state_dict().items()}
            cl = mode.startswith("amp")
            for k in sd:
                if cl and sd[k].dim() == 4:
                    sd[k] = sd[k].contiguous(memory_format=torch.channels_last)
            names = [k for k in sd if sd[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))]
            for k in names:
                sd[k].requires_grad_(True)
            opt = torch.optim.SGD([sd[k] for k in names], lr=1e-3, momentum=0.9, weight_decay=1e-4)
            W, b = sd["fusion_module.fc_out.weight"], sd["fusion_module.fc_out.bias"]

            def step():
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
                    a_in = spec.unsqueeze(1)
                    v_in = image
                    if cl:
                        a_in = a_in.contiguous(memory_format=torch.channels_last)
                    fa = torch.flatten(F.adaptive_avg_pool2d(resnet(sd, "audio_net.", a_in, False), 1), 1)
                    fv = resnet(sd, "visual_net.", v_in, True)
                    _, C, H, Wd = fv.shape
                    fv = torch.flatten(F.adaptive_avg_pool3d(fv.view(BATCH, -1, C, H, Wd).permute(0, 2, 1, 3, 4), 1), 1)
                tot = []
                for feat in (fa, fv):                                             # main.py:432-442 / 444-454
                    loss = F.cross_entropy(F.linear(feat.float(), W, b), label)
                    loss.backward()
                    opt.step()
                    opt.zero_grad()
                    tot.append(loss)
                return (tot[0] * 0.55 + tot[1] * 0.45).item(), tot[0].item(), tot[1].item()   # main.py:472-475

            for _ in range(warm):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[mode] = {"ms_per_step": ms, "samples_per_s": BATCH * 1e3 / ms}
            del sd, opt, net
            torch.cuda.empty_cache()
        except Exception as exc:                                                  # informational: never lose the headline
            res[mode] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    torch.backends.cudnn.allow_tf32 = prev
    res["what"] = ("the reference's alternating step under PyTorch eager on this GPU (cuDNN / ATen; B = 64, %d timed "
                   "steps after %d warm-up, CUDA events); GS hook as published = no-op" % (steps, warm))
    return res


def measure_tf32_peak(torch):
    """Measured dense TF32 rate of this GPU the way MEASURED_PEAKS.json measures bf16: torch.matmul (cuBLAS) on 8192^3 fp32
    operands with allow_tf32, best of 8, CUDA events. (BASELINE.md section 3 left the TF32 peak 'to do'; the roofline of the
    TF32 step is stated against this number instead of an assumed half of the bf16 peak.)"""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        x = torch.randn(n, n, device="cuda")
        y = torch.randn(n, n, device="cuda")
        best = float("inf")
        for _ in range(2):
            torch.matmul(x, y)
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(x, y); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return 2.0 * n ** 3 / best / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def tf32_leg(a):
    """The same K timed steps with TF32 operands (kind::tf32 tensor-core instructions, MLA_F16=0) in a child process:
    the operand format is fixed when the package is imported. Returns {value, ms_per_step} or {error}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--tf32-leg", "--steps", str(a.steps), "--warmup", str(a.warmup)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, MLA_F16="0"))
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as exc:
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def run_native(a, rank, world):
    import torch
    import mla_b200
    from mla_b200 import _lib, dist as mdist, encoder_engine, ops
    dev = torch.device("cuda", torch.cuda.current_device())
    args = make_args()
    mla_b200.setup_seed(0)
    model = mla_b200.ModuleHolder(mla_b200.AVClassifier(args).apply(mla_b200.weight_init).to(dev))
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin(force_projection=True)      # fire the projection (the published hook is a no-op)
    from mla_b200.main import SyntheticAVLoader
    host = SyntheticAVLoader(BATCH, 2, seed=1 + rank).batches                    # pinned host batches
    resident = [tuple(t.to(dev) for t in b) for b in host]                       # already in HBM
    pk = peaks()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def epoch(batches, n):
        return mla_b200.train_epoch(args, 0, model, dev, [batches[i % len(batches)] for i in range(n)], opt, sch,
                                    gs_plugin=gs, gs_flag=True, av_alpha=0.55)

    import contextlib, io
    quiet = contextlib.redirect_stdout(io.StringIO())
    with quiet:
        epoch(resident, a.warmup)
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count() + encoder_engine.GRAPH_LAUNCHES
    with quiet:
        t_dev = timed(lambda: epoch(resident, a.steps))
    launches = _lib.launch_count() + encoder_engine.GRAPH_LAUNCHES - l0     # eager launches + graph-replayed kernel nodes
    if a.tf32_leg:                          # child process of tf32_leg(): the device-resident timed region only
        print(json.dumps({"value": BATCH * world * a.steps / t_dev, "ms_per_step": 1e3 * t_dev / a.steps, "unit": "samples/s",
                          "dtype": "tf32" if not encoder_engine.USE_F16 else "f16",
                          "encoder_tflops": FLOP_PER_SAMPLE_STEP * BATCH * world * a.steps / t_dev / 1e12}), flush=True)
        return
    # data-parallel invariant: P, the head and both encoders are bit-identical on every rank
    dp_identical = None
    if world > 1:
        net_ = model.module
        cs = torch.tensor([mdist.params_checksum(gs.Pl), mdist.params_checksum(net_.fusion_module.fc_out.weight),
                           mdist.params_checksum(net_.audio_net._mla_flat), mdist.params_checksum(net_.visual_net._mla_flat),
                           gs.exp_count], dtype=torch.int64, device=dev)
        allc = [torch.zeros_like(cs) for _ in range(world)]
        torch.distributed.all_gather(allc, cs)
        dp_identical = all(torch.equal(allc[0], c) for c in allc)
        if not dp_identical:
            raise SystemExit("bench.py: data-parallel state differs across ranks: %s" % [c.tolist() for c in allc])
    # end to end: host (pinned) buffers in, H2D every step, losses read back to the host every step
    with quiet:
        epoch(host, min(2, a.warmup))

        # one epoch over K pinned host batches: H2D of every batch (prefetched one step ahead on a side stream)
        # and a device->host copy of every step's losses (args.step_loss_log) inside the timed region
        args.step_loss_log = True
        t_e2e = timed(lambda: epoch(host, a.steps))
        args.step_loss_log = False
        # evaluation with --dynamic (reported, not the headline); one untimed pass first (eval-mode graphs are captured)
        mla_b200.valid(args, model, dev, [resident[i % 2] for i in range(3)], gs_flag=True, av_alpha=0.55)
        t_eval = timed(lambda: mla_b200.valid(args, model, dev, [resident[i % 2] for i in range(a.steps)],
                                              gs_flag=True, av_alpha=0.55))
    clocks = sampler.stop() if rank == 0 else None
    # host cost of enqueueing one step (Python + ctypes + launches): the same launch sequence on a tiny batch, where
    # the GPU work is negligible — shows how far the step is from being launch-bound
    host_ms = None
    if rank == 0 and world == 1:
        tiny = SyntheticAVLoader(2, 2, seed=99, spec_hw=(65, 48), image_hw=(64, 64)).batches
        tiny = [tuple(t.to(dev) for t in b) for b in tiny]
        with quiet:
            epoch(tiny, 3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            epoch(tiny, 10)
            torch.cuda.synchronize()
            host_ms = (time.perf_counter() - t0) * 1e3 / 10
    # roofline leg: the SAME K steps once more with every convolution launch bracketed by CUDA events on the
    # launching stream (kept out of the headline timed region so that the event records cannot perturb it)
    conv = None
    from mla_b200 import basic_model
    overlap = basic_model.OVERLAP_ENCODERS
    basic_model.OVERLAP_ENCODERS = False     # one stream: per-launch durations must not include a co-running kernel
    overlap_w = encoder_engine._OVERLAP_WGRAD
    encoder_engine._OVERLAP_WGRAD = False
    if rank == 0:
        encoder_engine.CONV_TIMING = []
        with quiet:
            t_inst = timed(lambda: epoch(resident, a.steps)) if world == 1 else None
            if world > 1:
                epoch(resident, a.steps)
        torch.cuda.synchronize()
        recs = encoder_engine.CONV_TIMING
        encoder_engine.CONV_TIMING = None
        tot_t = sum(e0.elapsed_time(e1) for _, _, e0, e1 in recs) * 1e-3
        tot_f = sum(f for _, f, _, _ in recs)
        by = {}
        for kind, f, e0, e1 in recs:
            d = by.setdefault(kind, [0, 0.0, 0.0])
            d[0] += 1; d[1] += f; d[2] += e0.elapsed_time(e1) * 1e-3
        conv = {"launches": len(recs), "seconds": tot_t, "flops": tot_f,
                "by_kind": {k: {"launches": v[0], "tflops": v[1] / v[2] / 1e12, "ms_per_step": 1e3 * v[2] / a.steps,
                                "flops": v[1]} for k, v in by.items()},
                "instrumented_ms_per_step": None if t_inst is None else 1e3 * t_inst / a.steps}
    elif world > 1:
        with quiet:
            epoch(resident, a.steps)
    basic_model.OVERLAP_ENCODERS = overlap
    encoder_engine._OVERLAP_WGRAD = overlap_w
    if rank != 0:
        return
    samples = BATCH * world * a.steps
    value = samples / t_dev
    h2d = sum(t.numel() * t.element_size() for t in host[0][:3])
    out = {"metric": "MLA train samples/sec (CREMA-D AV, ResNet-18)", "value": value, "unit": "samples/s",
           "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_dev / a.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": ("f16" if encoder_engine.STEM_F16 else "f16+tf32") if encoder_engine.USE_F16 else "tf32",
           "dtype_note": ("tensor-core operands fp16 = TF32's 10-bit mantissa (forward: activations / weights; backward: "
                          "gradients multiplied by an exact power of two chosen per layer on the device, activations, "
                          "transposed filters), fp32 accumulation, fp32 conv outputs / BatchNorm / residual stream / SGD: "
                          "the arithmetic class of the reference under torch's default cudnn.allow_tf32=True "
                          "(tests/test_gpu_encoder_grad.py); value_tf32 = the same step on kind::tf32 instructions")
                         if encoder_engine.USE_F16 else "TF32 operands (kind::tf32), fp32 accumulation",
           "data": "synthetic",
           "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                      "encoder_backend": encoder_engine.BACKEND, "gs_projection": "fires (force_projection)",
                      "streams": ("audio / visual encoders on two CUDA streams" if overlap else "single stream") +
                                 ("; weight gradients on a third/fourth" if overlap_w else "") +
                                 ("; encoder launch sequences replayed as CUDA graphs" if encoder_engine.USE_GRAPHS else ""),
                      "l2": "2 alternating input batches (179 MB) + >1 GB of activations per step exceed the 126 MB L2; no explicit flush"},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": samples / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": 24, "ms_per_step": 1e3 * t_e2e / a.steps},
           "eval_samples_per_s": samples / t_eval, "host_enqueue_ms_per_step": host_ms,
           "dp_state_identical": dp_identical,
           "host_cpus_bound": (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None),
           "encoder_tflops": FLOP_PER_SAMPLE_STEP * samples / t_dev / 1e12}
    prof = profile_summary()
    if conv is not None and conv["seconds"] > 0:
        # dominant kernel of the step: conv_gemm_kernel (tcgen05 implicit GEMM; fprop + dgrad + wgrad launches).
        # achieved = algorithmic FLOPs (2*M*Cout*Cin*R*S per launch, stem with its true K=49*Cin) / CUDA-event time.
        # peak: the measured SUSTAINED bf16 figure (kernel timed inside a step) for the kind::f16 launches (fp16 / bf16
        # operands); HALF of it for the kind::tf32 launches (MEASURED_PEAKS.json has no TF32 figure; TF32 runs at half the
        # bf16 rate on sm_100: 1.1 vs 2.25 PFLOP/s nominal). The reported peak is the FLOP-weighted blend
        # sum(flops) / sum(flops_i / peak_i), i.e. frac = (time at peak) / (measured time).
        t_ideal = sum(v["flops"] / (pk["bf16_sustained"] * (1.0 if k.endswith("16") else 0.5) * 1e12)
                      for k, v in conv["by_kind"].items())
        peak = conv["flops"] / t_ideal / 1e12
        ach = conv["flops"] / conv["seconds"] / 1e12
        out["roofline"] = {"kernel": "conv16_persistent_kernel + conv_strip16_kernel (fprop16 / dgrad16) + conv_gemm_kernel (wgrad16) + stem_s2d kernels: tcgen05 implicit GEMM, TMA fed, kind::f16 (fp16 operands, scaled fp16 gradients)",
                           "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                           "traffic": prof.get("conv_traffic"),
                           "traffic_source": prof.get("conv_traffic_source", "ncu --set full capture (profiles/), not this run"),
                           "peak_source": "FLOP-weighted blend of bf16_tflops_sustained (f16 launches) and 0.5x (tf32 launches), "
                                          + pk["source"],
                           "launches_per_step": conv["launches"] / a.steps,
                           "avg_launch_us": 1e6 * conv["seconds"] / conv["launches"],
                           "flops_per_step": conv["flops"] / a.steps,
                           "share_of_step": (conv["seconds"] / a.steps) / (conv["instrumented_ms_per_step"] * 1e-3
                                                                           if conv["instrumented_ms_per_step"] else t_dev / a.steps),
                           "by_kind": {k: {kk: vv for kk, vv in v.items() if kk != "flops"} for k, v in conv["by_kind"].items()},
                           "instrumented_ms_per_step": conv["instrumented_ms_per_step"]}
    if not a.no_sweep and world == 1:
        # BASELINE.json configs[4]: D {512..2048} x B {64..4096} x C {6, 101} (x M {2, 3} for the fusion kernel). The headline
        # point of each kernel is its best bandwidth-regime point; every point (latency-bound ones labelled) is listed
        for key, fn, kern in (("gs", gs_sweep, "gs_project_kernel"), ("head", head_sweep, "head_rows_kernel + head_cols_kernel + head_reduce_kernel (C <= 16) / head_fwd_kernel + head_bwd_kernel"),
                              ("fusion", fusion_sweep, "fuse_eval_kernel")):
            try:
                pts = fn(torch, ops, pk)
            except Exception as exc:                                         # a failed sweep never costs the headline line
                out["roofline_" + key] = {"error": "%s: %s" % (type(exc).__name__, exc)}
                continue
            top = max((p for p in pts if p["regime"] != "fp32-compute"), key=lambda p: p["gbs"])
            out["roofline_" + key] = {"kernel": kern, "bound": "hbm", "achieved": top["gbs"], "peak": pk["hbm"],
                                      "unit": "GB/s", "frac": top["frac"],
                                      "traffic": prof.get("gs_traffic") if key == "gs" else None, "peak_source": pk["source"],
                                      "point": {k: v for k, v in top.items() if k not in ("gbs", "frac")},
                                      "latency_bound_points": sum(1 for p in pts if p["regime"] == "latency"),
                                      "points": len(pts),
                                      "timing": "CUDA events around ONE launch, mean of the central half of >= 24 launches (event timestamps tick every ~2 us); inputs cold: the launches "
                                                "rotate over `sets` independent argument sets (> 2x the 126 MB L2 in total when "
                                                "sets < 128), no L2 flush"}
            out[key + "_sweep"] = [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in p.items()} for p in pts]
        for key, leg in (("gs_beyond_sweep", lambda: [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in p.items()}
                                                      for p in gs_beyond_sweep(torch, ops, pk)]),
                         ("frame_producer", lambda: producer_leg(torch, pk))):
            try:                                                             # informational legs never cost the line
                out[key] = leg()
            except Exception as exc:
                out[key] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if not a.no_tf32_leg and world == 1 and encoder_engine.USE_F16:
        out["value_tf32"] = tf32_leg(a)
        try:
            tf32_peak = measure_tf32_peak(torch)
            out["value_tf32"]["cublas_tf32_tflops_measured"] = tf32_peak
            if "encoder_tflops" in out["value_tf32"]:
                out["value_tf32"]["frac_of_measured_tf32_peak"] = out["value_tf32"]["encoder_tflops"] / tf32_peak
        except Exception as exc:                                             # informational only
            out["value_tf32"]["cublas_tf32_tflops_measured"] = "error: %s" % exc
    if not a.no_eager and world == 1:
        out["eager_baseline"] = eager_baseline(torch)
    if not a.no_cpu_baseline and world == 1:
        try:
            out["cpu_baseline"] = cpu_baseline()
        except Exception as exc:
            out["cpu_baseline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if not a.no_extra and world == 1:
        out["other_workloads"] = other_workloads(torch)
    print(json.dumps(out), flush=True)


def other_workloads(torch, steps=3):
    """Informational, N = 1 only, outside the timed region of the headline metric: step time of the other parity
    configurations of BASELINE.json through the same public API (train_epoch), device-resident synthetic inputs.
    configs[2] = m3ae 'base' image-text pair, configs[3] = three-modality model; B = 64, S = 257 (+512 audio tokens)."""
    import argparse
    import contextlib
    import io
    import mla_b200
    from mla_b200.main import SyntheticTextImageLoader
    res = {}
    for name, modal3 in (("configs[2] m3ae image-text", False), ("configs[3] three-modality", True)):
        try:
            args = argparse.Namespace(dataset="IEMOCAP" if modal3 else "Food101", fusion_method="concat", modulation="Normal",
                                      gs_flag=True, dynamic=True, lorb="m3ae", modal3=modal3, clip=False)
            mla_b200.setup_seed(0)
            net = mla_b200.Modal3Classifier(args) if modal3 else mla_b200.M3AEClassifier(args)
            model = mla_b200.ModuleHolder(net.cuda())
            opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
            sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
            gs = mla_b200.GSPlugin()
            dev = torch.device("cuda")

            def run(n):
                loader = SyntheticTextImageLoader(BATCH, n, 1, n_classes=4 if modal3 else 101, audio_len=1024 if modal3 else 0)
                loader.batches = [tuple(t.cuda() if torch.is_tensor(t) else t for t in b) for b in loader.batches]
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                with contextlib.redirect_stdout(io.StringIO()):
                    mla_b200.train_epoch(args, 0, model, dev, loader, opt, sch, gs_plugin=gs, gs_flag=True, av_alpha=0.55)
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n
            run(2)
            ms = run(steps)
            res[name] = {"ms_per_step": ms, "samples_per_s": BATCH * 1e3 / ms, "batch": BATCH, "steps": steps,
                         "encoders": "m3ae base x2" + (" + CAV-MAE audio" if modal3 else "")}
            del model, net, opt
            torch.cuda.empty_cache()
        except Exception as exc:                                      # informational only: never lose the headline line
            res[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    return res


def start_watchdog(limit_s):
    """A wedged GPU must not turn into a silent hang: if the run has not finished after `limit_s` seconds, say so on
    stderr and leave with a non-zero exit code."""
    def fire():
        sys.stderr.write("bench.py watchdog: no result after %d s, aborting\n" % limit_s)
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(limit_s, fire)
    t.daemon = True
    t.start()
    return t


def main():
    a = parse()
    start_watchdog(900 if a.impl == "reference" else 600)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    from mla_b200 import dist as mdist
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the ONE JSON line
        mdist.init_from_env("nccl")
    else:
        torch.cuda.set_device(0)
    run_native(a, rank, world)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
