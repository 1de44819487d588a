#!/usr/bin/env python
"""Interleaved A/B timing of whole train steps (batch 8, 256x256, device-resident batch) in ONE process.

    python tools/ab_step.py base: epi0:FS_IN_EPILOGUE=0 pair0:FS_TC_PAIR=0 ...

Every argument is `name:ENV=VAL,ENV=VAL`; engine options are read from the environment when an engine is created
(FS_IN_EPILOGUE, FS_FAST_PREP, FS_TENSOR_PATH), FS_TC_PAIR is switched at run time.  All configurations are alive at
once and timed round-robin (ROUNDS x STEPS steps each, CUDA events), so clock / power / box drift hits all alike;
then each gets a per-kernel-class profile.  Prints a table and writes JSON to gpurun_out/ab_<names>.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from faststyle_b200 import _lib, synth  # noqa: E402
from faststyle_b200.engine import Engine, TFAdam, make_loss_config, pack_vgg, params_to_device  # noqa: E402
from faststyle_b200.tf_bundle import read_checkpoint  # noqa: E402

ROUNDS, STEPS = 4, 15


def main():
    specs = []
    for a in sys.argv[1:]:
        name, _, kv = a.partition(":")
        env = dict(x.split("=", 1) for x in kv.split(",") if x)
        specs.append((name, env))
    if not specs:
        specs = [("base", {})]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    raw = read_checkpoint(os.path.join(bench.GOLDEN, "starry_final.ckpt"))
    packed = pack_vgg(synth.synthetic_vgg_weights(7), dev)
    cfg = make_loss_config(bench.CONTENT_LAYERS, [1.0], bench.STYLE_LAYERS, [5.0] * 4, 0.0)
    style = bench.load_style_image()
    seng = Engine(1, style.shape[1], style.shape[2], vgg=True, style_layers=bench.STYLE_LAYERS, device=dev)
    tg = seng.vgg_grams(packed, style, bench.STYLE_LAYERS)
    torch.cuda.synchronize()
    del seng
    x = bench.synthetic_batch(0).to(dev)
    runs = []
    for name, env in specs:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        eng = Engine(bench.PER_GPU_BATCH, bench.HW, bench.HW, transform_bwd=True, vgg_bwd=True,
                     content_layers=bench.CONTENT_LAYERS, style_layers=bench.STYLE_LAYERS, device=dev)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        params = params_to_device(raw, dev)
        opt = TFAdam(params, 1e-3)
        grads = torch.empty_like(params)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        pair = int(env.get("FS_TC_PAIR", "1"))
        ew = int(env.get("FS_TC_EPI_WARPS", "16"))

        def step(eng=eng, params=params, opt=opt, grads=grads, losses=losses):
            eng.train_fwd_bwd(params, packed, x, cfg, tg, grads=grads, losses=losses)
            opt.step(grads)
        runs.append(dict(name=name, env=env, eng=eng, step=step, pair=pair, ew=ew, ms=[]))
    for r in runs:
        _lib.call("fs_set_tc_pair", r["pair"])
        _lib.call("fs_set_tc_epilogue_warps", r["ew"])
        for _ in range(5):
            r["step"]()
    torch.cuda.synchronize()
    for _ in range(ROUNDS):
        for r in runs:
            _lib.call("fs_set_tc_pair", r["pair"])
            _lib.call("fs_set_tc_epilogue_warps", r["ew"])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r["step"]()
            e0.record()
            for _ in range(STEPS):
                r["step"]()
            e1.record()
            torch.cuda.synchronize()
            r["ms"].append(e0.elapsed_time(e1) / STEPS)
    out = {}
    for r in runs:
        _lib.call("fs_set_tc_pair", r["pair"])
        _lib.call("fs_set_tc_epilogue_warps", r["ew"])
        prof = bench.live_kernel_profile(r["eng"], r["step"])
        out[r["name"]] = dict(env=r["env"], ms_per_step=r["ms"], best=min(r["ms"]), by_kernel_class=prof,
                              launches=sum(v["launches_per_step"] for v in prof.values()))
    _lib.call("fs_set_tc_pair", 1)
    _lib.call("fs_set_tc_epilogue_warps", 16)
    names = [r["name"] for r in runs]
    print("%-22s" % "" + "".join("%12s" % n for n in names))
    print("%-22s" % "ms/step (best)" + "".join("%12.3f" % out[n]["best"] for n in names))
    print("%-22s" % "ms/step (rounds)" + "".join("%12s" % ",".join("%.2f" % m for m in out[n]["ms_per_step"])[:12] for n in names))
    print("%-22s" % "bracketed launches" + "".join("%12d" % out[n]["launches"] for n in names))
    for c in bench.PROF_CATS:
        if any(c in out[n]["by_kernel_class"] for n in names):
            print("%-22s" % c + "".join("%12.3f" % out[n]["by_kernel_class"].get(c, {}).get("ms_per_step", 0.0) for n in names))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_%s.json" % "_".join(names)), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
