#!/usr/bin/env python
"""Run a few train steps (batch 8, 256x256, the bench.py workload) and exit - the target of the ncu
captures in profiles/ (a launch list with `--metrics gpu__time_duration.sum`, or `--set full -k regex:...`).

    python tools/step_once.py [steps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from faststyle_b200 import synth  # noqa: E402
from faststyle_b200.engine import Engine, TFAdam, make_loss_config, pack_vgg, params_to_device  # noqa: E402
from faststyle_b200.tf_bundle import read_checkpoint  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    params = params_to_device(read_checkpoint(os.path.join(bench.GOLDEN, "starry_final.ckpt")), dev)
    packed = pack_vgg(synth.synthetic_vgg_weights(7), dev)
    cfg = make_loss_config(bench.CONTENT_LAYERS, [1.0], bench.STYLE_LAYERS, [5.0] * 4, 0.0)
    style = bench.load_style_image()
    seng = Engine(1, style.shape[1], style.shape[2], vgg=True, style_layers=bench.STYLE_LAYERS, device=dev)
    tg = seng.vgg_grams(packed, style, bench.STYLE_LAYERS)
    torch.cuda.synchronize()
    del seng
    eng = Engine(bench.PER_GPU_BATCH, bench.HW, bench.HW, transform_bwd=True, vgg_bwd=True,
                 content_layers=bench.CONTENT_LAYERS, style_layers=bench.STYLE_LAYERS, device=dev)
    opt = TFAdam(params, 1e-3)
    x = bench.synthetic_batch(0).to(dev)
    grads = torch.empty_like(params)
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    from faststyle_b200 import _lib
    lib = _lib.load()
    n0 = lib.fs_launch_count()
    for _ in range(steps):
        eng.train_fwd_bwd(params, packed, x, cfg, tg, grads=grads, losses=losses)
        opt.step(grads)
    torch.cuda.synchronize()
    print("done", steps, "steps")
    print("launches_per_step", (lib.fs_launch_count() - n0) // steps)


if __name__ == "__main__":
    main()
