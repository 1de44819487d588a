#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: the 256x256 train step
(transform fwd + VGG16 perceptual loss + backward + TF-Adam), data-parallel.

    python bench.py --gpus N --steps K --warmup W           # our CUDA path
    python bench.py --impl reference ...                    # CPU restatement of the reference

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); weak scaling with
8 images per GPU (BASELINE.json configs[3] at N=8; configs[2]'s step at batch 8 for
N=1).  Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

PER_GPU_BATCH = 8
HW = 256
STYLE_LAYERS = ["conv1_2", "conv2_2", "conv3_3", "conv4_3"]
CONTENT_LAYERS = ["conv3_3"]
# algorithmic work per image of one train step, SURVEY.md 8(d): 60.97 GMAC
GFLOP_PER_IMAGE_TRAIN = 121.9
GFLOP_PER_IMAGE_FWD = 7.069
METRIC = "images/sec 256x256 train step (transform fwd + VGG16 perceptual loss + bwd + Adam)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_tflops=d.get("bf16_tflops", 1590.0),
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_batch(rank, n=PER_GPU_BATCH):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randint(0, 256, (n, HW, HW, 3), generator=g).float()


def load_style_image():
    import cv2
    img = cv2.cvtColor(cv2.imread(os.path.join(GOLDEN, "starry_night_crop.jpg")), cv2.COLOR_BGR2RGB)
    return img[None].astype(np.float32)


# ----------------------------------------------------------------------------- CPU (oracle) arm
def pick_cpu_threads():
    """Thread count for the CPU arm: the best of a short calibration over {8,16,32,64,all} cores
    (on a shared host `all` is often far slower than fewer threads - cgroup limits, SMT, NUMA)."""
    from oracle import ckpt as ockpt, restate as R
    params = ockpt.load(os.path.join(GOLDEN, "starry_final.ckpt"))
    x = synthetic_batch(0, 2).numpy()
    ncpu = os.cpu_count() or 1
    try:
        ncpu = min(ncpu, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu} | {ncpu})
    best, best_t = ncpu, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            R.create_net(x, params, "resize")
            t0 = time.perf_counter()
            R.create_net(x, params, "resize")
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_train_step_setup():
    from oracle import ckpt as ockpt, restate as R
    from faststyle_b200 import synth
    params = ockpt.load(os.path.join(GOLDEN, "starry_final.ckpt"))
    vggw = synth.synthetic_vgg_weights(7)
    style = load_style_image()[:, :256, :256]          # bounded: a 256x256 crop of the style image
    tg = R.style_target_grams(style, vggw, STYLE_LAYERS, torch.float32)
    p = {k: torch.from_numpy(v).clone() for k, v in params.items()}
    opt = R.TFAdam(p, 1e-3)

    def step(x):
        out = R.train_grads(x, p, vggw, tg, dtype=torch.float32, beta=0.0)
        opt.step(p, out["grads"])
        return float(out["loss"])
    return step


def run_reference(args, rank, world):
    """--impl reference: the reference's own TF1-CPU path cannot run here (TensorFlow
    is absent), so this arm times the CPU restatement of the same graph (oracle/) on all
    host threads.  Only rank 0 works."""
    if rank != 0:
        return
    pick_cpu_threads()
    step = cpu_train_step_setup()
    sample_b = 4
    x = synthetic_batch(0, sample_b).numpy()
    for _ in range(max(args.warmup, 1)):
        step(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(x)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample_b / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "train.py step 256x256 (CPU sample: batch %d per step)" % sample_b,
                   "per_gpu_batch": PER_GPU_BATCH, "parallelism": "cpu"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "torch-CPU restatement (not TF1), batch %d x %d steps" % (sample_b, args.steps)},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from faststyle_b200 import _lib, synth
    from faststyle_b200.engine import Engine, TFAdam, make_loss_config, pack_vgg, params_to_device
    from faststyle_b200.tf_bundle import read_checkpoint

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    params = params_to_device(read_checkpoint(os.path.join(GOLDEN, "starry_final.ckpt")), dev)
    packed = pack_vgg(synth.synthetic_vgg_weights(7), dev)
    cfg = make_loss_config(CONTENT_LAYERS, [1.0], STYLE_LAYERS, [5.0] * 4, 0.0)
    style = load_style_image()
    seng = Engine(1, style.shape[1], style.shape[2], vgg=True, style_layers=STYLE_LAYERS, device=dev)
    tgrams = seng.vgg_grams(packed, style, STYLE_LAYERS)
    torch.cuda.synchronize()
    del seng
    eng = Engine(PER_GPU_BATCH, HW, HW, transform_bwd=True, vgg_bwd=True, content_layers=CONTENT_LAYERS,
                 style_layers=STYLE_LAYERS, device=dev)
    opt = TFAdam(params, 1e-3)
    grads = torch.empty_like(params)
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    x_host = synthetic_batch(rank).pin_memory()
    x_dev = x_host.to(dev)
    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()

    def step_device():
        eng.train_fwd_bwd(params, packed, x_dev, cfg, tgrams, grads=grads, losses=losses)
        if world > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM)       # loss is a batch SUM (losses.py:32-37)
        opt.step(grads)

    # End-to-end step with the input copy of step i+1 overlapped with the compute of step i (the reference's queue
    # runners do the same on the host side): two device buffers, a copy stream, and events both ways - the compute
    # stream waits for "copy done", the copy stream waits for "last reader of this buffer done".
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [x_dev, torch.empty_like(x_dev)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def enqueue_copy(slot):
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            xbuf[slot].copy_(x_host, non_blocking=True)
            copied[slot].record(copy_stream)

    def step_e2e():
        main = torch.cuda.current_stream()
        slot = state["i"] & 1
        if not state["primed"]:
            consumed[0].record(main); consumed[1].record(main)
            enqueue_copy(slot)
            state["primed"] = True
        main.wait_event(copied[slot])
        eng.train_fwd_bwd(params, packed, xbuf[slot], cfg, tgrams, grads=grads, losses=losses)
        consumed[slot].record(main)
        enqueue_copy(slot ^ 1)                      # next step's input goes up while this step computes
        if world > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM)
        opt.step(grads)
        loss_host.copy_(losses, non_blocking=True)
        main.synchronize()
        state["i"] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample_clocks=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            # nvidia-smi needs a few hundred ms to deliver its first sample: keep the GPU under the
            # same load (untimed steps) until the sampler is running, then enter the timed region
            sampler.start()
            for _ in range(40):          # fixed count: every rank must issue the same collectives
                fn()
            barrier()
        n0 = lib.fs_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        n1 = lib.fs_launch_count()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), n1 - n0, clocks

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_e2e()
    ms, launches, clocks = timed(step_device, args.steps, sample_clocks=True)
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    imgs = PER_GPU_BATCH * world * args.steps
    value = imgs / (ms / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)

    # ---- extras (rank 0, N=1 semantics): forward-only throughput, roofline, CPU baseline
    extra = {}
    roof = None
    cpu_base = None
    if rank == 0:
        peaks = load_peaks()
        # whole-step tensor-pipe view (algorithmic FLOPs of one step / step time)
        step_tflops = GFLOP_PER_IMAGE_TRAIN * PER_GPU_BATCH / (ms / args.steps / 1e3) / 1e3
        # dominant kernel live: VGG conv3x3 64->64 @256^2 (conv1_2 shape) through the op C-ABI
        def step_local():           # rank-local (no collective): the other ranks are waiting at the barrier below
            eng.train_fwd_bwd(params, packed, x_dev, cfg, tgrams, grads=grads, losses=losses)
            opt.step(grads)
        prof = live_kernel_profile(eng, step_local)
        roof = dominant_kernel_roofline(prof, peaks, ms / args.steps)
        roof["step_tflops"] = step_tflops
        roof["step_frac_of_sustained"] = step_tflops / peaks["bf16_tflops_sustained"]
        # The legs below are additional lines of evidence; a failure in one of them (e.g. out of memory next to
        # another tenant) must not take the headline number with it.
        def leg(name, fn):
            try:
                return fn()
            except Exception as e:      # noqa: BLE001 - recorded in the JSON line, never silent
                torch.cuda.synchronize()
                return {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        # transform forward only, batch 32 (BASELINE.json configs[1])
        extra["transform_fwd_b32_images_per_s"] = leg("fwd", lambda: bench_forward(dev, params))
        if world == 1:       # BASELINE.json configs[4] (single-GPU Gatys optimisation at 1024x1024)
            extra["slow_style_1024"] = leg("slow_style", lambda: bench_slow_style(dev, packed, tgrams))
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = leg("cpu", cpu_baseline)
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "train.py step, per-GPU batch %d, %dx%d, starry ckpt weights, synthetic "
                                   "VGG16 weights, style starry_night_crop.jpg" % (PER_GPU_BATCH, HW, HW),
                       "global_batch": PER_GPU_BATCH * world, "parallelism": "dp%d" % world,
                       "l2": "per-step working set ~1.3 GB >> 126 MB L2, no explicit flush",
                       "precision": "fp32 storage; VGG conv1_2..conv4_3, the residual convs, the four stride-2 / "
                                    "resize convs (2x2 forms) incl. data + weight gradients, Gram fwd/bwd: split-bf16 x3 "
                                    "on tcgen05 (fp32-class, 16 mantissa bits), fp32 accumulate; 9x9 convs, conv1_1, "
                                    "InstanceNorm, pooling, losses, Adam: exact fp32 on CUDA cores"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s",
                    "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu_base,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


PROF_CATS = ["tc_vgg_conv_fwd", "tc_vgg_conv_dgrad", "tc_res_conv_fwd", "tc_res_conv_dgrad",
             "ffma_conv", "wgrad", "gram_fwd", "gram_bwd", "instnorm_stats", "instnorm_apply", "instnorm_bwd",
             "pointwise", "losses", "weight_prep", "tc_s2_conv_fwd", "tc_s2_conv_dgrad"]


def live_kernel_profile(eng, step_fn, steps=3):
    """Per-kernel-class device time inside real train steps: CUDA events recorded by the engine
    around every GEMM-class launch on the launching stream (fs_engine_profile)."""
    import ctypes as C
    from faststyle_b200 import _lib
    n = len(PROF_CATS)
    _lib.call("fs_engine_profile", eng._h, 1)
    for _ in range(steps):
        step_fn()
    torch.cuda.synchronize()
    ms = (C.c_float * n)(); fl = (C.c_double * n)(); cnt = (C.c_int * n)()
    _lib.call("fs_engine_profile_read", eng._h, n, ms, fl, cnt)
    _lib.call("fs_engine_profile", eng._h, 0)
    out = {}
    for i, name in enumerate(PROF_CATS):
        if cnt[i]:
            out[name] = {"ms_per_step": ms[i] / steps, "launches_per_step": cnt[i] // steps,
                         "tflops": fl[i] / (ms[i] / 1e3) / 1e12 if ms[i] > 0 else None,
                         "gflop_per_step": fl[i] / steps / 1e9}
    return out


def dominant_kernel_roofline(prof, peaks, step_ms):
    """Dominant kernel = the tcgen05 3x3 convolution (VGG fwd + dgrad + residual convs):
    achieved = algorithmic FLOPs of those launches (2*MACs; the three split-bf16 passes are NOT
    multiplied in) / their summed event time."""
    names = [k for k in ("tc_vgg_conv_fwd", "tc_vgg_conv_dgrad", "tc_res_conv_fwd", "tc_res_conv_dgrad",
                         "tc_s2_conv_fwd", "tc_s2_conv_dgrad") if k in prof]
    if names:
        ms = sum(prof[k]["ms_per_step"] for k in names)
        gf = sum(prof[k]["gflop_per_step"] for k in names)
        launches = sum(prof[k]["launches_per_step"] for k in names)
        kernel, precision = "conv3x3_tc_kernel (tcgen05, TMA slabs, TMEM accumulators)", "split-bf16 x3 (hi*hi+hi*lo+lo*hi), fp32 accumulate"
    else:
        ms = prof["ffma_conv"]["ms_per_step"]; gf = prof["ffma_conv"]["gflop_per_step"]
        launches = prof["ffma_conv"]["launches_per_step"]
        kernel, precision = "igemm_kernel (FFMA)", "fp32"
    achieved = gf / ms                      # GFLOP / ms == TFLOP/s
    return {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops_sustained"],
            # DRAM bytes per launch from the committed `ncu --set full` capture (profiles/r01p_ncu_tc_kernels.md),
            # averaged over the 30 captured launches of this kernel (content pass, transform convs, main VGG pass);
            # per-launch values are tabulated there (e.g. conv1_2: 352 MB moved = 134 MB split input + 217 MB
            # output planes - no re-reads from HBM).
            "traffic": 59.1e6 if names else None, "traffic_unit": "bytes/launch (ncu, mean of 30 captured launches)",
            "peak_source": peaks["source"] +
            " cuBLAS bf16 sustained (kernel timed inside a long step)",
            "kernel": kernel, "precision": precision, "kernel_ms_per_step": ms, "launches_per_step": launches,
            "share_of_step": ms / step_ms,
            "mma_issue_frac": 3.0 * achieved / peaks["bf16_tflops_sustained"] if names else None,
            "by_kernel_class": prof}


def bench_forward(dev, params):
    from faststyle_b200.engine import Engine
    B = 32
    eng = Engine(B, HW, HW, transform=True, device=dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (B, HW, HW, 3), generator=g).float().to(dev)
    for _ in range(3):
        eng.transform_forward(params, x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 5
    e0.record()
    for _ in range(reps):
        eng.transform_forward(params, x)
    e1.record()
    torch.cuda.synchronize()
    return B * reps / (e0.elapsed_time(e1) / 1e3)


def bench_slow_style(dev, packed, tgrams):
    """BASELINE.json configs[4]: slow_style.py 1024x1024 Gatys iteration (VGG16 fwd + Gram + losses + pixel gradient
    + TF-Adam on the image, lr 10, beta 1e-4) - iterations/s and algorithmic TFLOP/s (1.2356 TFLOP per iteration)."""
    from faststyle_b200.engine import Engine, TFAdam, make_loss_config
    S = 1024
    eng = Engine(1, S, S, vgg_bwd=True, content_layers=CONTENT_LAYERS, style_layers=STYLE_LAYERS, device=dev)
    cfg = make_loss_config(CONTENT_LAYERS, [1.0], STYLE_LAYERS, [5.0] * 4, 1e-4)
    g = torch.Generator().manual_seed(0)
    content = torch.randint(0, 256, (1, S, S, 3), generator=g).float().to(dev)
    g2 = torch.Generator().manual_seed(2)
    x = (torch.rand((1, S, S, 3), generator=g2) * 255.0).to(dev).contiguous()
    eng.set_content_targets(packed, content, cfg)
    opt = TFAdam(x.view(-1), 10.0)

    def it():
        _, grad = eng.perceptual_loss(packed, x, cfg, tgrams)
        opt.step(grad.view(-1))
    for _ in range(3):
        it()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 10
    e0.record()
    for _ in range(reps):
        it()
    e1.record()
    torch.cuda.synchronize()
    ips = reps / (e0.elapsed_time(e1) / 1e3)
    return {"iters_per_s": ips, "ms_per_iter": 1e3 / ips, "tflops_algorithmic": 1.2356 * ips}


def cpu_baseline():
    pick_cpu_threads()
    step = cpu_train_step_setup()
    b = 4
    x = synthetic_batch(0, b).numpy()
    step(x)
    n = 12                                 # ~10 s of CPU work on the GPU box's host cores
    t0 = time.perf_counter()
    for _ in range(n):
        step(x)
    dt = (time.perf_counter() - t0) / n
    return {"value": b / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "torch-CPU restatement of the train step (not TF1): batch %d, %d steps after 1 warm-up" % (b, n)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
