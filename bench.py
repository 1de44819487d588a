#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: the 256x256 train step
(transform fwd + VGG16 perceptual loss + backward + TF-Adam), data-parallel.

    python bench.py --gpus N --steps K --warmup W           # our CUDA path
    python bench.py --impl reference ...                    # CPU restatement of the reference

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); weak scaling with
8 images per GPU (BASELINE.json configs[3] at N=8; configs[2]'s step at batch 8 for
N=1).  Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for definitions.

Both timed legs go through `faststyle_b200.trainer.Trainer.step`, the call train.py makes:
`value` with the batch resident on the device, `e2e` with the batch in pinned host memory (host->device copy of
every step's input and a device->host read of every step's loss inside the timed region).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

PER_GPU_BATCH = 8
STRONG_GLOBAL_BATCH = 64          # strong scaling with global 64
HW = 256
STYLE_LAYERS = ["conv1_2", "conv2_2", "conv3_3", "conv4_3"]
CONTENT_LAYERS = ["conv3_3"]
# algorithmic work per image of one train step, SURVEY.md 8(d): 60.97 GMAC
GFLOP_PER_IMAGE_TRAIN = 121.9
GFLOP_PER_IMAGE_FWD = 7.069
METRIC = "images/sec 256x256 train step (transform fwd + VGG16 perceptual loss + bwd + Adam)"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_tflops=d.get("bf16_tflops", 1590.0),
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def csrc_sha16():
    """Hash of the CUDA sources: ties an ncu capture in profiles/ to the build it was taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "faststyle_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


KERNEL_SOURCES = ("conv3x3_tc.cu", "tc.cuh", "tc_ptx.cuh", "common.cuh")


def kernel_sha16():
    """Hash of the dominant kernel's own sources (conv3x3_tc_kernel and the headers it includes): an ncu capture of
    that kernel stays valid while other kernels of the library change."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "faststyle_b200", "csrc")
    for f in KERNEL_SOURCES:
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every ~5 ms DURING the timed region (a thread of this
    process: nvidia-smi's own loop delivers 2-3 samples in a 0.1-0.3 s region)."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, gpu_index):
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: NVML indexes physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.sm, self.power, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, attr in self.REASONS:
                    if r & getattr(nv, attr):
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join(2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "how": "NVML polled every 5 ms inside the timed region"}


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
    """The reference's train step (train.py:245-280) restated on the CPU: same starry weights, same synthetic VGG
    weights, the FULL style image - the configuration of the GPU arm."""
    from oracle import ckpt as ockpt, restate as R
    from faststyle_b200 import synth
    params = ockpt.load(os.path.join(GOLDEN, "starry_final.ckpt"))
    vggw = synth.synthetic_vgg_weights(7)
    tg = R.style_target_grams(load_style_image(), vggw, STYLE_LAYERS, torch.float32)
    p = {k: torch.from_numpy(v).clone() for k, v in params.items()}
    opt = R.TFAdam(p, 1e-3)

    def step(x):
        out = R.train_grads(x, p, vggw, tg, dtype=torch.float32, beta=0.0)
        opt.step(p, out["grads"])
        return float(out["loss"])
    return step


def workload_string(batch):
    return ("train.py step, per-GPU batch %d, %dx%d, starry ckpt weights, synthetic VGG16 weights, style "
            "starry_night_crop.jpg" % (batch, HW, HW))


def run_reference(args, rank, world):
    """--impl reference: the reference's own TF1-CPU path cannot run here (TensorFlow
    is absent), so this arm times the CPU restatement of the same graph (oracle/) on all
    host threads, on the GPU arm's configuration (batch 8 per step).  Only rank 0 works."""
    if rank != 0:
        return
    cores = pick_cpu_threads()
    step = cpu_train_step_setup()
    x = synthetic_batch(0, PER_GPU_BATCH).numpy()
    warm = max(args.warmup, 1)
    for _ in range(warm):
        step(x)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step(x)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    val = PER_GPU_BATCH / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_string(PER_GPU_BATCH), "global_batch": PER_GPU_BATCH,
                   "parallelism": "cpu (one host, %d threads)" % cores,
                   "note": "CPU restatement of the reference graph (torch-CPU fp32); TensorFlow 1.x is not installable here"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "torch-CPU restatement (not TF1), batch %d x %d steps after %d warm-ups; median "
                                   "step %.1f ms" % (PER_GPU_BATCH, args.steps, warm, statistics.median(times) * 1e3)},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def make_trainer(dev, batch, pg, params, vggw, style):
    from faststyle_b200.trainer import Trainer
    return Trainer(params, vggw, style, batch, (HW, HW), CONTENT_LAYERS, STYLE_LAYERS, [1.0], [5.0] * 4, 0.0, 1e-3,
                   device=dev, process_group=pg)


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from faststyle_b200 import _lib, synth
    from faststyle_b200.tf_bundle import read_checkpoint

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    lib = _lib.load()

    params = read_checkpoint(os.path.join(GOLDEN, "starry_final.ckpt"))
    vggw = synth.synthetic_vgg_weights(7)
    style = load_style_image()
    trainer = make_trainer(dev, PER_GPU_BATCH, pg, params, vggw, style)
    x_host = synthetic_batch(rank).pin_memory()

    def step_device():                                  # batch resident in HBM
        trainer.step(None, fetch_losses=False)

    def step_e2e():                                     # pinned host batch up + loss down, every step
        trainer.step(x_host, fetch_losses="lag")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample_clocks=False, finish=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            for _ in range(10):          # fixed count on every rank: get the GPU to its loaded clock state first
                fn()
            barrier()
            sampler.start()
        def launches():      # kernels issued through the library + kernels executed by the step-graph replays
            return lib.fs_launch_count() + trainer.replayed_launches
        n0 = launches()
        e0.record()
        for _ in range(steps):
            fn()
        if finish:
            finish()
        e1.record()
        barrier()
        n1 = launches()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), n1 - n0, clocks

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_e2e()
    trainer.flush_losses()
    ms, launches, clocks = timed(step_device, args.steps, sample_clocks=True)
    ms_e2e, _, _ = timed(step_e2e, args.steps, finish=trainer.flush_losses)
    imgs = PER_GPU_BATCH * world * args.steps
    value = imgs / (ms / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)

    # replicas must hold identical parameters after the run (same reduced gradient, same Adam step everywhere)
    replicas_identical = None
    if world > 1:
        cs = torch.tensor([trainer.param_checksum()], dtype=torch.float64, device=dev)
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        replicas_identical = all(float(c.item()) == float(allcs[0].item()) for c in allcs)
        assert replicas_identical, "data-parallel replicas diverged: parameter checksums %s" % [float(c) for c in allcs]

    # ---- live per-kernel-class profile (rank 0), taken directly after the timed legs - in the same clock / thermal
    # state - and before the heavier legs below.  Forward/backward only and no optimiser step: the replicas (and the
    # device-side step counter the fused exchange tags its flags with) must stay in lock-step for the next leg.
    prof = None
    if rank == 0:
        eng0 = trainer.engine
        prof = live_kernel_profile(eng0, lambda: eng0.train_fwd_bwd(
            trainer.params, trainer.packed, trainer._current, trainer.cfg, trainer.target_grams,
            grads=trainer.grads, losses=trainer.losses))

    # ---- strong-scaling leg (every rank takes part): global batch 64 split over the ranks
    strong = None
    try:
        if STRONG_GLOBAL_BATCH % world == 0:
            lb = STRONG_GLOBAL_BATCH // world
            tr2 = trainer if lb == PER_GPU_BATCH else make_trainer(dev, lb, pg, params, vggw, style)
            xs = synthetic_batch(rank, lb).pin_memory()
            for _ in range(3):
                tr2.step(xs, fetch_losses=False)
            k = max(5, args.steps // 4)
            ms_s, _, _ = timed(lambda: tr2.step(None, fetch_losses=False), k)
            strong = {"global_batch": STRONG_GLOBAL_BATCH, "per_gpu_batch": lb, "steps": k, "ms_per_step": ms_s / k,
                      "images_per_s": STRONG_GLOBAL_BATCH * k / (ms_s / 1e3), "scaling": "strong"}
            if tr2 is not trainer:
                del tr2
                torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001 - recorded, never silent (all ranks fail alike: same shapes everywhere)
        torch.cuda.synchronize()
        strong = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    # ---- extras (rank 0, rank-local: the other ranks wait at the barrier below)
    extra = {}
    roof = None
    cpu_base = None
    if rank == 0:
        peaks = load_peaks()
        step_ms = ms / args.steps
        eng = trainer.engine

        roof = dominant_kernel_roofline(prof, peaks, step_ms, clocks)
        step_tflops = GFLOP_PER_IMAGE_TRAIN * PER_GPU_BATCH / (step_ms / 1e3) / 1e3
        roof["whole_step"] = {"tflops_algorithmic": step_tflops, "frac_of_burst": step_tflops / peaks["bf16_tflops"],
                              "frac_of_sustained": step_tflops / peaks["bf16_tflops_sustained"]}

        # The legs below are additional lines of evidence; a failure in one of them (e.g. out of memory next to
        # another tenant) must not take the headline number with it.
        def leg(fn):
            try:
                return fn()
            except Exception as e:      # noqa: BLE001 - recorded in the JSON line, never silent
                torch.cuda.synchronize()
                return {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        # BASELINE.json configs[1]: transform forward only, batch 32
        extra["transform_fwd_b32"] = leg(lambda: bench_forward(dev, params, peaks))
        # BASELINE.json configs[2]: the train step at batch 4
        extra["train_step_b4"] = leg(lambda: bench_train_b4(dev, params, vggw, style))
        if world == 1:       # BASELINE.json configs[4] (single-GPU Gatys optimisation at 1024x1024)
            extra["slow_style_1024"] = leg(lambda: bench_slow_style(dev, trainer.packed, trainer.target_grams))
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = leg(cpu_baseline)
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(PER_GPU_BATCH),
                       "global_batch": PER_GPU_BATCH * world, "parallelism": "dp%d" % world,
                       "l2": "per-step working set ~1.3 GB >> 126 MB L2, no explicit flush",
                       "precision": "fp32 storage; VGG conv1_2..conv4_3, the residual convs, the four stride-2 / "
                                    "resize convs (2x2 forms) incl. data + weight gradients, the 9x9 convs (fwd + data "
                                    "gradient), Gram fwd/bwd: split-bf16 x3 on tcgen05 (fp32-class, 16 mantissa bits), "
                                    "fp32 accumulate; conv1_1 fwd: three-way bf16 split, six products on tcgen05 "
                                    "(24 mantissa bits); 9x9 weight gradients, conv1_1 data gradient, InstanceNorm, "
                                    "losses, Adam: exact fp32 on CUDA cores",
                       "dp_exchange": (None if world == 1 else
                                       "fs_dp_allreduce_adam: one kernel, gradient all-reduce over NVLink peer loads + Adam"
                                       if trainer.peer is not None else "NCCL all_reduce + Adam kernel"),
                       "api": "faststyle_b200.trainer.Trainer.step (the call train.py makes)",
                       "step_issue": ("CUDA-graph replay of the step's launches (captured once per input buffer)"
                                      if trainer._use_graph and trainer._graphs else "one launch per kernel")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s",
                    "h2d_bytes_per_step": int(trainer.h2d_bytes_per_step), "d2h_bytes_per_step": int(trainer.d2h_bytes_per_step),
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "pinned host batch -> device on a copy stream (double-buffered, overlaps the previous step); "
                           "each step's 4 loss scalars read back to pinned host memory one step later (all inside the "
                           "timed region)"},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu_base,
            "strong_scaling": strong,
        }
        if replicas_identical is not None:
            line["replicas_identical"] = replicas_identical
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


PROF_CATS = ["tc_vgg_conv_fwd", "tc_vgg_conv_dgrad", "tc_res_conv_fwd", "tc_res_conv_dgrad",
             "ffma_conv", "wgrad", "gram_fwd", "gram_bwd", "instnorm_stats", "instnorm_apply", "instnorm_bwd",
             "pointwise", "losses", "weight_prep", "tc_s2_conv_fwd", "tc_s2_conv_dgrad", "tc_9x9_conv_fwd",
             "tc_9x9_conv_dgrad", "tc_conv1_1_fwd"]
TC_CONV_CATS = ("tc_vgg_conv_fwd", "tc_vgg_conv_dgrad", "tc_res_conv_fwd", "tc_res_conv_dgrad",
                "tc_s2_conv_fwd", "tc_s2_conv_dgrad", "tc_9x9_conv_fwd", "tc_9x9_conv_dgrad")


def live_kernel_profile(eng, step_fn, steps=3):
    """Per-kernel-class device time inside real steps: CUDA events recorded by the engine
    around every GEMM-class launch on the launching stream (fs_engine_profile)."""
    import ctypes as C
    from faststyle_b200 import _lib
    n = len(PROF_CATS)
    _lib.call("fs_engine_profile", eng._h, 1)
    for _ in range(steps):
        step_fn()
    torch.cuda.synchronize()
    ms = (C.c_float * n)(); fl = (C.c_double * n)(); cnt = (C.c_int * n)(); by = (C.c_double * n)()
    _lib.call("fs_engine_profile_bytes", eng._h, n, by)
    _lib.call("fs_engine_profile_read", eng._h, n, ms, fl, cnt)
    _lib.call("fs_engine_profile", eng._h, 0)
    out = {}
    for i, name in enumerate(PROF_CATS):
        if cnt[i]:
            out[name] = {"ms_per_step": ms[i] / steps, "launches_per_step": cnt[i] // steps,
                         "tflops": fl[i] / (ms[i] / 1e3) / 1e12 if ms[i] > 0 else None,
                         "gflop_per_step": fl[i] / steps / 1e9}
            if by[i] > 0:
                out[name]["algorithmic_mb_per_step"] = by[i] / steps / 1e6
    return out


def load_ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of THIS build
    (tools/ncu_capture.sh -> tools/ncu_summarize.py traffic -> profiles/ncu_traffic.json).  None when there is no
    capture; `build_matches` says whether the capture's source hash equals the sources benched now."""
    if not os.path.exists(TRAFFIC_FILE):
        return None
    try:
        d = json.load(open(TRAFFIC_FILE))
    except Exception:
        return None
    d["build_matches"] = d.get("csrc_sha16") == csrc_sha16()                 # every CUDA source identical
    d["kernel_source_matches"] = d.get("kernel_sha16") == kernel_sha16()     # the captured kernel's sources identical
    return d


def dominant_kernel_roofline(prof, peaks, step_ms, clocks):
    """Dominant kernel = the tcgen05 convolution (VGG fwd + dgrad, residual, stride-2 / resize forms):
    achieved = algorithmic FLOPs of those launches (2*MACs; the three split-bf16 passes are NOT
    multiplied in) / their summed CUDA-event time inside real steps."""
    names = [k for k in TC_CONV_CATS if k in prof]
    if names:
        ms = sum(prof[k]["ms_per_step"] for k in names)
        gf = sum(prof[k]["gflop_per_step"] for k in names)
        mb = sum(prof[k].get("algorithmic_mb_per_step", 0.0) for k in names)
        launches = sum(prof[k]["launches_per_step"] for k in names)
        kernel, precision = "conv3x3_tc_kernel (tcgen05, TMA slabs, TMEM accumulators)", "split-bf16 x3 (hi*hi+hi*lo+lo*hi), fp32 accumulate"
    else:
        ms = prof["ffma_conv"]["ms_per_step"]; gf = prof["ffma_conv"]["gflop_per_step"]; mb = 0.0
        launches = prof["ffma_conv"]["launches_per_step"]
        kernel, precision = "igemm_kernel (FFMA)", "fp32"
    achieved = gf / ms                      # GFLOP / ms == TFLOP/s
    # Which measured peak applies: the burst figure when the timed region ran at (near) maximum SM clock with no
    # power cap - a sub-second region does - the sustained one when the clocks show the power-capped state.
    burst = True
    if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz"):
        # (a sw_power_cap flag seen in some samples of a region whose MEDIAN clock is the maximum is noted in `clocks`,
        # but does not make the sustained peak - measured at a 1.2 GHz median - the honest denominator)
        burst = clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"]
    peak = peaks["bf16_tflops"] if burst else peaks["bf16_tflops_sustained"]
    traffic = load_ncu_traffic() if names else None
    out = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
           "frac_of_burst": achieved / peaks["bf16_tflops"],
           "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"],
           "peak_source": "%s cuBLAS bf16 %s (MEASURED_PEAKS.json): timed region ran at %s" %
                          (peaks["source"], "burst" if burst else "sustained",
                           "max SM clock, no power cap" if burst else "power-capped clocks"),
           "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
           "traffic_source": ({k: traffic.get(k) for k in ("file", "launches", "build_matches", "kernel_source_matches",
                                                             "csrc_sha16", "kernel_sha16", "unit")}
                              if traffic else "no ncu capture of this build committed (profiles/ncu_traffic.json absent)"),
           "algorithmic_bytes_per_launch": mb * 1e6 / launches if launches and mb else None,
           "kernel": kernel, "precision": precision, "kernel_ms_per_step": ms, "launches_per_step": launches,
           "share_of_step": ms / step_ms,
           "explain": {"mma_issue_frac_of_peak": 3.0 * achieved / peak if names else None,
                       "note": "every logical MAC is three bf16 MMAs (hi*hi + hi*lo + lo*hi): the algorithmic ceiling of "
                               "this kernel is peak/3; `frac` counts algorithmic FLOPs only"},
           "by_kernel_class": prof}
    return out


def bench_forward(dev, params, peaks):
    """BASELINE.json configs[1]: transform-net inference at batch 32, 256x256.  Device-resident throughput through
    Engine.transform_forward (weights prepared once), end-to-end throughput through BatchStylizer (uint8 host
    batch in, uint8 host batch out, pipelined copies), and the per-class profile of the forward pass."""
    from faststyle_b200.engine import Engine, params_to_device
    from faststyle_b200.stream import BatchStylizer
    B = 32
    eng = Engine(B, HW, HW, transform=True, device=dev)
    eng.set_frozen_weights(True)
    pflat = params_to_device(params, dev)
    g = torch.Generator().manual_seed(0)
    xu8 = torch.randint(0, 256, (B, HW, HW, 3), generator=g, dtype=torch.uint8)
    x = xu8.float().to(dev)
    for _ in range(3):
        eng.transform_forward(pflat, x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 10
    e0.record()
    for _ in range(reps):
        eng.transform_forward(pflat, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ips = B / (ms / 1e3)
    prof = live_kernel_profile(eng, lambda: eng.transform_forward(pflat, x))
    del eng
    st = BatchStylizer(params, B, HW, HW, device=dev)
    xp = xu8.pin_memory()
    st.submit(xp); st.submit(xp); st.fetch(copy=False)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        st.submit(xp)
        st.fetch(copy=False)
    st.fetch(copy=False)
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = t0.elapsed_time(t1) / (reps + 1)
    tfl = GFLOP_PER_IMAGE_FWD * B / (ms / 1e3) / 1e3
    top = max(prof.items(), key=lambda kv: kv[1]["ms_per_step"])
    return {"images_per_s": ips, "ms_per_batch": ms, "batch": B,
            "e2e": {"value": B / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(st.h2d_bytes),
                    "d2h_bytes_per_step": int(st.d2h_bytes), "ms_per_batch": ms_e2e,
                    "api": "faststyle_b200.stream.BatchStylizer.submit/fetch (uint8 host batch in, uint8 host batch out)"},
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": tfl / peaks["bf16_tflops"],
                         "note": "whole forward pass, 7.069 GFLOP/image in the reference's formulation (SURVEY 8d); "
                                 "top class by time: %s (%.3f ms)" % (top[0], top[1]["ms_per_step"]),
                         "by_kernel_class": prof}}


def bench_train_b4(dev, params, vggw, style):
    """BASELINE.json configs[2]: train.py's default batch of 4 on one GPU (device-resident batch)."""
    tr = make_trainer(dev, 4, None, params, vggw, style)
    x = synthetic_batch(0, 4).pin_memory()
    for _ in range(3):
        tr.step(x, fetch_losses=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 10
    e0.record()
    for _ in range(reps):
        tr.step(None, fetch_losses=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        tr.step(x, fetch_losses="lag")
    tr.flush_losses()
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / reps
    return {"batch": 4, "images_per_s": 4 / (ms / 1e3), "ms_per_step": ms, "e2e_images_per_s": 4 / (ms2 / 1e3)}


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
    """3 warm-ups + >= 5 timed steps, median - on the GPU arm's configuration (batch 8)."""
    cores = pick_cpu_threads()
    step = cpu_train_step_setup()
    b = PER_GPU_BATCH
    x = synthetic_batch(0, b).numpy()
    for _ in range(3):
        step(x)
    times = []
    for _ in range(7):                     # ~10-15 s of CPU work on the GPU box's host cores
        t0 = time.perf_counter()
        step(x)
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    return {"value": b / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "torch-CPU restatement of the train step (not TF1): batch %d, median of %d steps after 3 warm-ups"
                      % (b, len(times))}


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
