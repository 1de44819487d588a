"""Data-parallel semantics on CPU (gloo, world_size 2): the gradient of the reference's loss
on a batch equals the SUM of the shard gradients (losses are batch sums, InstanceNorm is per
sample - SURVEY 7.2), so ONE all-reduce(SUM) of the flat gradient reproduces the 1-rank step,
and identical TF-Adam updates keep the replicas in lock-step.  The gradients here come from the
oracle (the CUDA path cannot run without a GPU); the wiring (flat layout, SUM reduction,
replica-identical Adam) is what train.py / Trainer use."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from faststyle_b200 import synth
    from faststyle_b200.layout import flatten_transform, unflatten_transform
    from oracle import restate as R
    params = synth.init_transform_params(seed=1)           # identical on every rank
    vggw = synth.synthetic_vgg_weights(7)
    rng = np.random.RandomState(0)
    style = rng.randint(0, 256, (1, 32, 32, 3)).astype(np.float32)
    tg = R.style_target_grams(style, vggw, ("conv1_2", "conv2_2"), torch.float64)
    full = rng.randint(0, 256, (2, 48, 48, 3)).astype(np.float32)
    kw = dict(style_layers=("conv1_2", "conv2_2"), style_weights=(5.0, 5.0), content_layers=("conv2_1",),
              content_weights=(1.0,), beta=1e-4, dtype=torch.float64)
    shard = full[rank:rank + 1]
    out = R.train_grads(shard, params, vggw, tg, **kw)
    flat = torch.from_numpy(flatten_transform({k: v.numpy() for k, v in out["grads"].items()}).astype(np.float64))
    loss = out["loss"].clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    # replica-identical TF-Adam step on the reduced gradient
    p = {k: torch.from_numpy(v).double() for k, v in params.items()}
    opt = R.TFAdam(p, 1e-3)
    g = {k: torch.from_numpy(v).double() for k, v in unflatten_transform(flat.numpy()).items()}
    opt.step(p, g)
    if rank == 0:
        ref = R.train_grads(full, params, vggw, tg, **kw)
        ref_flat = flatten_transform({k: v.numpy() for k, v in ref["grads"].items()}).astype(np.float64)
        q.put((float((flat - torch.from_numpy(ref_flat)).abs().max() / np.abs(ref_flat).max()),
               abs(float(loss) - float(ref["loss"])) / float(ref["loss"])))
    after = torch.from_numpy(flatten_transform({k: v.numpy() for k, v in p.items()}))
    gathered = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    if rank == 0:
        q.put(float((gathered[0] - gathered[1]).abs().max()))
    dist.destroy_process_group()


def test_two_rank_sum_allreduce_equals_single_rank_step():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    grad_err, loss_err = q.get(timeout=10)
    replica_diff = q.get(timeout=10)
    assert grad_err < 1e-6 and loss_err < 1e-12      # the flat layout is float32
    assert replica_diff == 0.0
