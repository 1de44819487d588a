#!/usr/bin/env python
"""Multi-GPU check of the fused data-parallel step (run under torchrun on >= 2 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Two trainers per rank on the SAME batches: one on fs_dp_allreduce_adam (gradient all-reduce over NVLink peer loads
fused with Adam), one on NCCL all_reduce + the Adam kernel.  Checks after every step that (1) all ranks hold
bit-identical parameters on the fused path, (2) the fused kernel by itself reproduces all_reduce + Adam on known
gradients to <= 1e-6, and whole steps agree with the NCCL path to <= 1e-5 of the largest parameter from the same state,
(3) the summed loss scalars agree; then times both (CUDA events, max over ranks).  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from faststyle_b200 import synth  # noqa: E402
from faststyle_b200.tf_bundle import read_checkpoint  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
    params = read_checkpoint(os.path.join(bench.GOLDEN, "starry_final.ckpt"))
    vggw = synth.synthetic_vgg_weights(7)
    style = bench.load_style_image()
    B = 2
    os.environ["FS_DP_FUSED"] = "1"
    fused = bench.make_trainer(dev, B, pg, params, vggw, style)
    os.environ["FS_DP_FUSED"] = "0"
    plain = bench.make_trainer(dev, B, pg, params, vggw, style)
    assert plain.peer is None
    out = {"world": world, "fused_available": fused.peer is not None}
    if fused.peer is None:
        if rank == 0:
            print(json.dumps(out), flush=True)
        dist.destroy_process_group()
        return
    # ---- (A) the kernel by itself: known per-rank gradients (wide dynamic range), three steps, against
    #      NCCL all_reduce(SUM) + the single-GPU Adam kernel from the same state
    from faststyle_b200.engine import TFAdam
    from faststyle_b200.layout import TRANSFORM_NPARAMS
    from faststyle_b200.peer import PeerExchange
    n = TRANSFORM_NPARAMS
    px = PeerExchange(n, 4, dev, pg)
    g0 = torch.Generator(device="cpu").manual_seed(1234)
    p0 = torch.randn(n, generator=g0).to(dev)
    opt_a, opt_b = TFAdam(p0.clone(), 1e-3), TFAdam(p0.clone(), 1e-3)
    gr = torch.Generator(device="cpu").manual_seed(77 + rank)
    worst_k = 0.0
    for t in range(3):
        g = (torch.randn(n, generator=gr) * torch.pow(10.0, torch.rand(n, generator=gr) * 6 - 6)).to(dev)
        e = torch.rand(4, generator=gr).to(dev)
        par = t & 1
        px.grads(par).copy_(g)
        px.extra(par).copy_(e)
        px.allreduce_adam(opt_a, par, t + 1)
        G, E = g.clone(), e.clone()
        dist.all_reduce(G)
        dist.all_reduce(E)
        opt_b.step(G)
        torch.cuda.synchronize()
        px.check()
        for a, b in ((opt_a.p, opt_b.p), (opt_a.m, opt_b.m), (opt_a.v, opt_b.v), (px.extra_sum, E)):
            worst_k = max(worst_k, float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30))
        assert int(opt_a.step_counter.item()) == t + 1
    assert worst_k <= 1e-6, "fs_dp_allreduce_adam vs all_reduce + fs_adam_step: %g" % worst_k
    out["kernel_maxdiff_rel"] = worst_k
    px.close()

    # ---- (B) whole training steps on the same batches; the NCCL trainer is re-synchronised to the fused one before
    #      every step (two runs of the same step differ at the 1e-8 level by themselves - fp64 atomics in arrival order
    #      - and Adam turns a flipped sign of a near-zero gradient into a full-size update, so free-running copies drift)
    rng = np.random.RandomState(100 + rank)
    worst_p, worst_l = 0.0, 0.0
    for step in range(4):
        x = rng.randint(0, 256, (B, bench.HW, bench.HW, 3)).astype(np.float32)
        for dst, src in ((plain.params, fused.params), (plain.opt.m, fused.opt.m), (plain.opt.v, fused.opt.v),
                         (plain.opt.step_counter, fused.opt.step_counter)):
            dst.copy_(src)
        lf = fused.step(x, fetch_losses=True)
        lp = plain.step(x, fetch_losses=True)
        fused.peer.check()
        cs = torch.tensor([fused.param_checksum()], dtype=torch.float64, device=dev)
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        assert all(float(c) == float(allcs[0]) for c in allcs), "fused path: replicas diverged at step %d: %s" % (step, allcs)
        d = float((fused.params - plain.params).abs().max()) / float(plain.params.abs().max())
        worst_p = max(worst_p, d)
        worst_l = max(worst_l, float(np.max(np.abs(lf - lp) / np.maximum(np.abs(lp), 1e-30))))
        assert d <= 1e-5 and worst_l <= 1e-6, "fused vs NCCL: params %g, losses %g" % (d, worst_l)
    out.update({"param_maxdiff_rel": worst_p, "loss_maxdiff_rel": worst_l, "replicas_identical": True})

    def timed(tr, steps=30):
        for _ in range(5):
            tr.step(None, fetch_losses=False)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.step(None, fetch_losses=False)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
    for _ in range(2):
        out["ms_per_step_fused"] = timed(fused)
        out["ms_per_step_nccl"] = timed(plain)
    fused.peer.check()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
