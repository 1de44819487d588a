#!/usr/bin/env python
"""Does replaying the whole train step as ONE CUDA graph beat issuing its ~150 launches?  (batch 8, 256x256)

    python tools/graph_step.py

Captures train_fwd_bwd + Adam (same buffers every step) with torch.cuda.CUDAGraph after a warm-up and times eager
issue against graph replay, interleaved.  Prints ms per step for both and the host time per eager step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from faststyle_b200 import synth  # noqa: E402
from faststyle_b200.engine import Engine, TFAdam, make_loss_config, pack_vgg, params_to_device  # noqa: E402
from faststyle_b200.tf_bundle import read_checkpoint  # noqa: E402


def main():
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
    B = int(os.environ.get("FS_GRAPH_BATCH", bench.PER_GPU_BATCH))
    eng = Engine(B, bench.HW, bench.HW, transform_bwd=True, vgg_bwd=True, content_layers=bench.CONTENT_LAYERS,
                 style_layers=bench.STYLE_LAYERS, device=dev)
    opt = TFAdam(params, 1e-3)
    x = bench.synthetic_batch(0)[:B].to(dev)
    grads = torch.empty_like(params)
    losses = torch.empty(4, dtype=torch.float32, device=dev)

    def step():
        eng.train_fwd_bwd(params, packed, x, cfg, tg, grads=grads, losses=losses)
        opt.step(grads)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step()
        torch.cuda.synchronize()

        def timed(fn, n=30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn()
            e0.record()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            host = (time.perf_counter() - t0) / n * 1e3
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n, host
        res = {"eager": [], "graph": []}
        for _ in range(4):
            res["eager"].append(timed(step))
            res["graph"].append(timed(g.replay))
    print("batch", B)
    for k, v in res.items():
        print("%-6s ms/step (device) %s   host ms/step %s" % (k, ["%.3f" % a for a, _ in v], ["%.3f" % b for _, b in v]))


if __name__ == "__main__":
    main()
