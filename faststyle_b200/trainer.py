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
        # ONE flat exchange buffer per step: [424102 gradients | pad | content, style, tv, total].  A single
        # all-reduce(SUM) carries the gradients and the loss scalars for logging.
        off = (TRANSFORM_NPARAMS + 7) // 8 * 8
        self._flat = torch.zeros(off + 8, dtype=torch.float32, device=self.device)
        self.grads = self._flat[:TRANSFORM_NPARAMS]
        self.losses = self._flat[off:off + 4]
        # Multi-GPU on one node: the all-reduce and Adam run as ONE kernel over NVLink peer memory
        # (fs_dp_allreduce_adam, faststyle_b200/peer.py) when every rank can map its peers' exchange buffers;
        # otherwise (FS_DP_FUSED=0, gloo / CPU harness, IPC unavailable) one NCCL all-reduce + the Adam kernel.
        self.peer = self._setup_peer_exchange() if self.world > 1 else None
        # The device part of a step (forward/backward composite + exchange + Adam: ~150 launches, ~3 ms of host time
        # at batch 8 against ~4 ms on the device) is captured ONCE per (input buffer, exchange-buffer parity) as a CUDA
        # graph and replayed: the host cost of a step drops to microseconds (measured: 4.00 -> 3.92 ms per step at batch
        # 8, 2.54 -> 2.44 at batch 4).  Not with the NCCL fallback (its collective stays outside any capture);
        # FS_STEP_GRAPH=0 issues the launches every step.
        import os
        self._use_graph = os.environ.get("FS_STEP_GRAPH", "1") != "0" and (self.world == 1 or self.peer is not None)
        self._graphs = {}
        self._eager_steps = 0
        self.replayed_launches = 0           # kernels executed by graph replays (fs_launch_count sees only eager launches)
        # input double buffering: the host->device copy of batch i+1 runs on a copy stream while step i computes
        self.x_dev = [torch.empty((self.batch_size, H, W, 3), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.x_host = [torch.empty((self.batch_size, H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._copied = [torch.cuda.Event(), torch.cuda.Event()]       # copy stream: batch landed in x_dev[slot]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]     # compute stream: last reader of x_dev[slot] done
        self._slot = 0
        self._current = self.x_dev[0]
        self.loss_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._loss_slot = 0
        self._loss_pending = None
        self.global_step = 0
        self.h2d_bytes_per_step = self.x_dev[0].numel() * 4
        self.d2h_bytes_per_step = 16

    def _setup_peer_exchange(self):
        import os
        import torch.distributed as dist
        ok, px = 1, None
        if os.environ.get("FS_DP_FUSED", "1") == "0" or dist.get_backend(self.pg) != "nccl":
            ok = 0
        else:
            try:
                from .peer import PeerExchange
                px = PeerExchange(TRANSFORM_NPARAMS, 4, self.device, self.pg)
            except Exception as e:                       # noqa: BLE001 - any failure means "use NCCL", on ALL ranks
                print("faststyle_b200: peer-memory exchange unavailable (%s); using NCCL all-reduce" % (e,), flush=True)
                ok, px = 0, None
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)      # the choice must be the same on every rank
        return px if int(flag.item()) == 1 else None

    # ------------------------------------------------------------------ input staging
    def _stage(self, batch):
        """Bring ``batch`` to the device through the double buffer; returns the device tensor the step reads."""
        if isinstance(batch, torch.Tensor) and batch.is_cuda:
            return batch                                   # GpuPreprocessor output: already resident
        main = torch.cuda.current_stream(self.device)
        slot = self._slot
        self._slot ^= 1
        if isinstance(batch, np.ndarray) or not batch.is_pinned():
            if isinstance(batch, np.ndarray):
                batch = torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32))
            self._copied[slot].synchronize()               # the previous copy out of this pinned slot has finished
            self.x_host[slot].copy_(batch)
            batch = self.x_host[slot]
        self.copy_stream.wait_event(self._consumed[slot])
        with torch.cuda.stream(self.copy_stream):
            self.x_dev[slot].copy_(batch, non_blocking=True)
            self._copied[slot].record(self.copy_stream)
        main.wait_event(self._copied[slot])
        self._consumed_slot = slot
        return self.x_dev[slot]

    def step(self, batch=None, fetch_losses=True):
        """One optimisation step on this rank's shard ``batch`` (host NHWC float32 array / pinned tensor / device
        tensor; None re-uses the batch already on the device).

        ``fetch_losses``: True - return [content, style, tv, total] summed over ALL ranks (one 16-byte
        device->host copy + a stream synchronise); "lag" - return the losses of the PREVIOUS step that asked for
        them (no host stall: the read of step i overlaps step i+1; ``flush_losses()`` drains the last one);
        False - return None.

        Data-parallel runs must call this the same number of times on every rank: use
        ``datapipe.next_batch_collective`` to end the input on all ranks at the same step."""
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            self._consumed_slot = None
            if batch is not None:
                self._current = self._stage(batch)
            if self._use_graph and self._eager_steps >= 2 and self._replay_or_capture():
                if self._consumed_slot is not None:
                    self._consumed[self._consumed_slot].record(main)
            elif self.peer is not None:
                self._eager_steps += 1
                par = self.global_step & 1
                self.grads, local_losses = self.peer.grads(par), self.peer.extra(par)
                self.engine.train_fwd_bwd(self.params, self.packed, self._current, self.cfg, self.target_grams,
                                          grads=self.grads, losses=local_losses)
                if self._consumed_slot is not None:
                    self._consumed[self._consumed_slot].record(main)
                self.peer.allreduce_adam(self.opt, par, self.global_step + 1)
                self.losses = self.peer.extra_sum
            else:
                self._eager_steps += 1
                self.engine.train_fwd_bwd(self.params, self.packed, self._current, self.cfg, self.target_grams,
                                          grads=self.grads, losses=self.losses)
                if self._consumed_slot is not None:
                    self._consumed[self._consumed_slot].record(main)
                if self.world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.pg)
                self.opt.step(self.grads)
            self.global_step += 1
            if not fetch_losses:
                return None
            ls = self._loss_slot
            self._loss_slot ^= 1
            self.loss_host[ls].copy_(self.losses, non_blocking=True)
            self._loss_ready[ls].record(main)
            if fetch_losses == "lag":
                prev, self._loss_pending = self._loss_pending, ls
                if prev is None:
                    return None
                self._loss_ready[prev].synchronize()
                return self.loss_host[prev].clone().numpy()
            self._loss_ready[ls].synchronize()
            return self.loss_host[ls].clone().numpy()

    def _device_step(self, par, tag):
        """forward/backward composite + (exchange +) Adam on the current stream, for buffer parity ``par``."""
        if self.peer is not None:
            self.engine.train_fwd_bwd(self.params, self.packed, self._current, self.cfg, self.target_grams,
                                      grads=self.peer.grads(par), losses=self.peer.extra(par))
            self.peer.allreduce_adam(self.opt, par, tag)
        else:
            self.engine.train_fwd_bwd(self.params, self.packed, self._current, self.cfg, self.target_grams,
                                      grads=self.grads, losses=self.losses)
            self.opt.step(self.grads)

    def _replay_or_capture(self) -> bool:
        """Replay the CUDA graph of this step's (input buffer, parity), capturing it first if needed.  Returns False
        (and turns graphs off) when the capture fails - the caller then issues the step eagerly."""
        from . import _lib
        par = self.global_step & 1 if self.peer is not None else 0
        key = (self._current.data_ptr(), par)
        entry = self._graphs.get(key)
        if entry is None:
            try:
                lib = _lib.load()
                g = torch.cuda.CUDAGraph()
                n0 = lib.fs_launch_count()
                with torch.cuda.graph(g):
                    self._device_step(par, 0)                # tag 0: the step number is read on the device
                entry = (g, int(lib.fs_launch_count() - n0))
                self._graphs[key] = entry
            except Exception as e:                           # noqa: BLE001 - graphs are an optimisation only
                print("faststyle_b200: CUDA-graph capture of the train step failed (%s); issuing launches" % (e,), flush=True)
                self._use_graph = False
                return False
        if self.peer is not None:
            self.grads, self.losses = self.peer.grads(par), self.peer.extra_sum
        entry[0].replay()
        self.replayed_launches += entry[1]
        return True

    def flush_losses(self):
        """Losses of the last ``fetch_losses="lag"`` step (None if there is none pending)."""
        prev, self._loss_pending = self._loss_pending, None
        if prev is None:
            return None
        self._loss_ready[prev].synchronize()
        return self.loss_host[prev].clone().numpy()

    def param_checksum(self) -> float:
        """fp64 sum of the flat parameter buffer: equal on every rank of a data-parallel run."""
        return float(self.params.double().sum().item())

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
