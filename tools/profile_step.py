#!/usr/bin/env python
"""Per-launch device time of one train step (batch 8, 256x256): the engine brackets every GEMM-class /
element-wise launch with CUDA events on the launching stream (fs_engine_profile_records).

    python tools/profile_step.py [out.md]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from faststyle_b200 import _lib, synth  # noqa: E402
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
    eng = Engine(bench.PER_GPU_BATCH, bench.HW, bench.HW, transform_bwd=True, vgg_bwd=True,
                 content_layers=bench.CONTENT_LAYERS, style_layers=bench.STYLE_LAYERS, device=dev)
    opt = TFAdam(params, 1e-3)
    x = bench.synthetic_batch(0).to(dev)
    grads = torch.empty_like(params)
    losses = torch.empty(4, dtype=torch.float32, device=dev)

    def step():
        eng.train_fwd_bwd(params, packed, x, cfg, tg, grads=grads, losses=losses)
        opt.step(grads)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    reps = 5
    acc = None
    for _ in range(reps):
        _lib.call("fs_engine_profile", eng._h, 1)
        step()
        torch.cuda.synchronize()
        mx = 1024
        cat = (C.c_int * mx)(); ms = (C.c_float * mx)(); fl = (C.c_double * mx)(); cnt = C.c_int()
        _lib.call("fs_engine_profile_records", eng._h, mx, cat, ms, fl, C.byref(cnt))
        n = cnt.value
        rec = [(cat[i], ms[i], fl[i]) for i in range(n)]
        nn = len(bench.PROF_CATS)
        _lib.call("fs_engine_profile_read", eng._h, nn, (C.c_float * nn)(), (C.c_double * nn)(), (C.c_int * nn)())
        _lib.call("fs_engine_profile", eng._h, 0)
        if acc is None:
            acc = [[c, m, f] for c, m, f in rec]
        else:
            assert len(acc) == len(rec)
            for a, r in zip(acc, rec):
                a[1] += r[1]
    lines = ["| # | class | us | GFLOP | TFLOP/s |", "|---|---|---|---|---|"]
    tot = 0.0
    for i, (c, m, f) in enumerate(acc):
        us = m / reps * 1e3
        tot += us
        lines.append("| %d | %s | %.1f | %.2f | %s |" % (i, bench.PROF_CATS[c], us, f / 1e9,
                                                        "%.0f" % (f / (us * 1e-6) / 1e12) if f > 0 else ""))
    lines.append("")
    lines.append("total %.1f us over %d bracketed launches (mean of %d steps; event brackets add ~2 us each)" % (tot, len(acc), reps))
    out = "\n".join(lines)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as fh:
            fh.write(out + "\n")
    print(out)


if __name__ == "__main__":
    main()
