"""Training-step driver shared by train.py and bench.py: flat parameters + fused TF-Adam on
the device, the fused forward/backward composite, and data-parallel gradient all-reduce.

Multi-GPU: one process per GPU, the batch is sharded across ranks, every rank
holds a replica of the 1.7 MB transform parameters and the frozen VGG weights, and ONE
all-reduce(SUM) of the flat 424102-float gradient follows the backward pass (the reference's
losses are SUMS over the batch, losses.py:32-37,63-64, so the global gradient is the sum of
the shard gradients; InstanceNorm statistics are per sample, so nothing else is exchanged)."""
from __future__ import annotations

import numpy as np
import torch

from .engine import Engine, TFAdam, make_loss_config, pack_vgg, params_to_device
from .layout import TRANSFORM_NPARAMS, unflatten_transform


class Trainer:
    def __init__(self, params: dict, vgg_weights: dict, style_img, batch_size, preprocess_size,
                 content_layers, style_layers, content_weights, style_weights, beta, learn_rate,
                 device="cuda:0", process_group=None, upsample_method="resize"):
        self.device = torch.device(device)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        H, W = int(preprocess_size[0]), int(preprocess_size[1])
        self.batch_size = int(batch_size)
        self.upsample_method = upsample_method
        self.params = params_to_device(params, self.device, upsample_method)
        self.packed = pack_vgg(vgg_weights, self.device)
        self.cfg = make_loss_config(content_layers, content_weights, style_layers, style_weights, beta)
        # target Grams of the style image (train.py:143-151)
        style = np.asarray(style_img, np.float32)
        seng = Engine(1, style.shape[1], style.shape[2], vgg=True, style_layers=style_layers, device=self.device)
        self.target_grams = seng.vgg_grams(self.packed, style, style_layers)
        torch.cuda.synchronize(self.device)
        del seng
        self.engine = Engine(self.batch_size, H, W, transform_bwd=True, vgg_bwd=True,
                             content_layers=content_layers, style_layers=style_layers, device=self.device,
                             deconv=upsample_method == "deconv")
        self.opt = TFAdam(self.params, learn_rate)
        self.grads = torch.empty(TRANSFORM_NPARAMS, dtype=torch.float32, device=self.device)
        self.losses = torch.empty(4, dtype=torch.float32, device=self.device)
        self.x_dev = torch.empty((self.batch_size, H, W, 3), dtype=torch.float32, device=self.device)
        self.loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
        self.global_step = 0

    def step(self, batch=None, fetch_losses=True):
        """One optimisation step on this rank's shard ``batch`` (host NHWC float32 array / pinned
        tensor; None re-uses the batch already on the device).  Returns [content, style, tv, total]
        summed over ALL ranks when ``fetch_losses`` (one small device->host copy), else None."""
        if batch is not None:
            if isinstance(batch, np.ndarray):
                batch = torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32))
            self.x_dev.copy_(batch, non_blocking=True)
        self.engine.train_fwd_bwd(self.params, self.packed, self.x_dev, self.cfg, self.target_grams,
                                  grads=self.grads, losses=self.losses)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.pg)
            if fetch_losses:
                dist.all_reduce(self.losses, op=dist.ReduceOp.SUM, group=self.pg)
        self.opt.step(self.grads)
        self.global_step += 1
        if not fetch_losses:
            return None
        self.loss_host.copy_(self.losses, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.loss_host.clone().numpy()

    def variables(self) -> dict:
        """Current transform-net variables as {tf_name: ndarray} (for tf.train.Saver-style saving)."""
        return unflatten_transform(self.params.detach().cpu().numpy(), self.upsample_method)

    def optimizer_slots(self) -> dict:
        """Adam slot variables under TF's names (<var>/Adam, <var>/Adam_1, beta*_power)."""
        m = unflatten_transform(self.opt.m.detach().cpu().numpy(), self.upsample_method)
        v = unflatten_transform(self.opt.v.detach().cpu().numpy(), self.upsample_method)
        out = {}
        for k in m:
            out[k + "/Adam"] = m[k]
            out[k + "/Adam_1"] = v[k]
        t = self.global_step
        out["beta1_power"] = np.array(self.opt.b1 ** (t + 1), np.float32)
        out["beta2_power"] = np.array(self.opt.b2 ** (t + 1), np.float32)
        return out
