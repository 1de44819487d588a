#!/usr/bin/env python
"""Multi-GPU check of the fused data-parallel step (run under torchrun on >= 2 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Two trainers per rank on the SAME batches: one on fs_dp_allreduce_adam (gradient all-reduce over NVLink peer loads
fused with Adam), one on NCCL all_reduce + the Adam kernel.  Checks after every step that (1) all ranks hold
bit-identical parameters on the fused path, (2) the fused path agrees with the NCCL path to <= 1e-6 of the largest
parameter (two runs of the same step differ at the 1e-8 level by themselves: the InstanceNorm statistics are summed
with fp64 atomics in arrival order), (3) the summed loss scalars agree; then times both (CUDA events, max over ranks).  Prints one JSON line on rank 0."""
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
    rng = np.random.RandomState(100 + rank)
    worst_p, worst_l = 0.0, 0.0
    for step in range(4):
        x = rng.randint(0, 256, (B, bench.HW, bench.HW, 3)).astype(np.float32)
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
        assert d <= 1e-6 and worst_l <= 1e-6, "fused vs NCCL: params %g, losses %g" % (d, worst_l)
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
