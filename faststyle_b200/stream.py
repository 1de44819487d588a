"""Frame-by-frame stylisation (the reference's stylize_webcam.py:76-103 loop body) on one B200.

One fixed-shape plan, pinned host staging buffers, and the whole per-frame device sequence
(uint8 -> fp32, weight preparation, the 16-conv transform network, fp32 -> uint8 with the
reference's channel swap) captured ONCE into a CUDA graph and replayed per frame: the ~70 short
kernels of a batch-1 forward are launch-latency-bound, so a single graph launch per frame is what
keeps latency at the sum of the kernels.  ``FS_STREAM_GRAPH=0`` replays them as plain stream
launches instead (same kernels, same results).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .engine import Engine, _require_cuda, params_to_device, ptr, stream_ptr


class FrameStylizer:
    """stylize(frame_uint8[H, W, 3]) -> uint8 [OH, OW, 3].

    ``swap_rb`` reproduces the reference loop exactly: the BGR camera frame is fed un-swapped and the
    result goes through ``cv2.cvtColor(..., COLOR_BGR2RGB)`` (stylize_webcam.py:82-90); the float output
    is truncated like ``ndarray.astype(np.uint8)`` (:89)."""

    def __init__(self, params: dict, height: int, width: int, upsample_method: str = "resize",
                 device="cuda:0", swap_rb: bool = True, use_graph: bool | None = None):
        _require_cuda()
        self.device = torch.device(device)
        self.H, self.W = int(height), int(width)
        self.swap_rb = bool(swap_rb)
        self.engine = Engine(1, self.H, self.W, transform=True, device=self.device,
                             deconv=upsample_method == "deconv")
        self.engine.set_frozen_weights(True)       # one model per stylizer: the captured graph holds no weight prep
        self.OH, self.OW = self.engine.OH, self.engine.OW
        self.params = params_to_device(params, self.device, upsample_method)
        self.in_host = torch.empty((self.H, self.W, 3), dtype=torch.uint8).pin_memory()
        self.out_host = torch.empty((self.OH, self.OW, 3), dtype=torch.uint8).pin_memory()
        self.in_dev = torch.empty((self.H, self.W, 3), dtype=torch.uint8, device=self.device)
        self.x_dev = torch.empty((1, self.H, self.W, 3), dtype=torch.float32, device=self.device)
        self.y_dev = torch.empty((1, self.OH, self.OW, 3), dtype=torch.float32, device=self.device)
        self.out_dev = torch.empty((self.OH, self.OW, 3), dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        if use_graph is None:
            use_graph = os.environ.get("FS_STREAM_GRAPH", "1") != "0"
        self.graph = None
        self.frames = 0
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._enqueue()                       # warm-up: first-launch attribute setup happens outside capture
            self.stream.synchronize()
            if use_graph:
                g = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(g, stream=self.stream):
                        self._enqueue()
                    self.graph = g
                except Exception as e:            # same kernels as plain launches; only the launch vehicle differs
                    import warnings
                    warnings.warn("CUDA-graph capture of the frame pipeline failed (%s); using stream launches" % e)
                    torch.cuda.synchronize(self.device)
                    self.graph = None

    def _enqueue(self):
        s = stream_ptr()
        _lib.call("fs_frame_u8_to_f32", ptr(self.in_dev), ptr(self.x_dev), C.c_longlong(self.in_dev.numel()), s)
        _lib.call("fs_transform_forward", self.engine._h, ptr(self.params), ptr(self.x_dev), ptr(self.y_dev), s)
        _lib.call("fs_frame_f32_to_u8", ptr(self.y_dev), ptr(self.out_dev), C.c_longlong(self.OH * self.OW),
                  1 if self.swap_rb else 0, s)

    def stylize(self, frame: np.ndarray) -> np.ndarray:
        if frame.dtype != np.uint8 or tuple(frame.shape) != (self.H, self.W, 3):
            raise _lib.FsError("expected a uint8 frame of shape %s, got %s %s"
                               % ((self.H, self.W, 3), frame.dtype, tuple(frame.shape)))
        self.in_host.numpy()[...] = frame
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.in_dev.copy_(self.in_host, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._enqueue()
            self.out_host.copy_(self.out_dev, non_blocking=True)
            self.stream.synchronize()
        self.frames += 1
        return self.out_host.numpy().copy()


class BatchStylizer:
    """Pipelined batch inference (BASELINE config 2: transform-net forward at batch B): uint8 images in host
    memory -> stylised uint8 images in host memory, the way stylize_image.py handles one image
    (uint8 in, float cast at the feed, cv2.imwrite's round-and-saturate on the way out, stylize_image.py:58-80).

    Two slots of pinned staging buffers and three streams: the host->device copy of batch i+1 and the device->host
    copy of batch i-1 run beside the compute of batch i.  ``submit(batch)`` enqueues a batch and returns at once;
    ``fetch()`` blocks for the oldest outstanding result; ``stylize(batch)`` = submit + fetch."""

    def __init__(self, params: dict, batch: int, height: int, width: int, upsample_method: str = "resize",
                 device="cuda:0"):
        _require_cuda()
        self.device = torch.device(device)
        self.B, self.H, self.W = int(batch), int(height), int(width)
        self.engine = Engine(self.B, self.H, self.W, transform=True, device=self.device,
                             deconv=upsample_method == "deconv")
        self.engine.set_frozen_weights(True)
        self.OH, self.OW = self.engine.OH, self.engine.OW
        self.params = params_to_device(params, self.device, upsample_method)
        ishape, oshape = (self.B, self.H, self.W, 3), (self.B, self.OH, self.OW, 3)
        self.in_host = [torch.empty(ishape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.out_host = [torch.empty(oshape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.in_dev = [torch.empty(ishape, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.out_dev = [torch.empty(oshape, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.x_dev = torch.empty(ishape, dtype=torch.float32, device=self.device)
        self.y_dev = torch.empty(oshape, dtype=torch.float32, device=self.device)
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_cmp = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]        # in_dev[slot] filled
        self.ev_in_free = [torch.cuda.Event() for _ in range(2)]   # in_dev[slot] read by the compute stream
        self.ev_cmp = [torch.cuda.Event() for _ in range(2)]       # out_dev[slot] written
        self.ev_out = [torch.cuda.Event() for _ in range(2)]       # out_host[slot] filled
        self._submitted = 0
        self._fetched = 0
        self.h2d_bytes = self.in_host[0].numel()
        self.d2h_bytes = self.out_host[0].numel()

    def submit(self, batch):
        """batch: uint8 [B,H,W,3] numpy array or (pinned) torch tensor."""
        if self._submitted - self._fetched >= 2:
            raise _lib.FsError("BatchStylizer: two batches are already in flight; fetch() one first")
        slot = self._submitted & 1
        if isinstance(batch, np.ndarray):
            batch = torch.from_numpy(np.ascontiguousarray(batch))
        if batch.dtype != torch.uint8 or tuple(batch.shape) != (self.B, self.H, self.W, 3):
            raise _lib.FsError("expected a uint8 batch of shape %s, got %s %s"
                               % ((self.B, self.H, self.W, 3), batch.dtype, tuple(batch.shape)))
        with torch.cuda.device(self.device):
            if not batch.is_pinned():
                self.ev_in[slot].synchronize()              # the last copy out of this staging slot is done
                self.in_host[slot].copy_(batch)
                batch = self.in_host[slot]
            self.s_in.wait_event(self.ev_in_free[slot])
            with torch.cuda.stream(self.s_in):
                self.in_dev[slot].copy_(batch, non_blocking=True)
                self.ev_in[slot].record(self.s_in)
            self.s_cmp.wait_event(self.ev_in[slot])
            self.s_cmp.wait_event(self.ev_out[slot])        # out_dev[slot] has left for the host
            with torch.cuda.stream(self.s_cmp):
                s = stream_ptr()
                _lib.call("fs_frame_u8_to_f32", ptr(self.in_dev[slot]), ptr(self.x_dev),
                          C.c_longlong(self.in_dev[slot].numel()), s)
                self.ev_in_free[slot].record(self.s_cmp)
                _lib.call("fs_transform_forward", self.engine._h, ptr(self.params), ptr(self.x_dev), ptr(self.y_dev), s)
                _lib.call("fs_frame_f32_to_u8", ptr(self.y_dev), ptr(self.out_dev[slot]),
                          C.c_longlong(self.B * self.OH * self.OW), 2, s)
                self.ev_cmp[slot].record(self.s_cmp)
            self.s_out.wait_event(self.ev_cmp[slot])
            with torch.cuda.stream(self.s_out):
                self.out_host[slot].copy_(self.out_dev[slot], non_blocking=True)
                self.ev_out[slot].record(self.s_out)
        self._submitted += 1

    def fetch(self, copy=True):
        if self._fetched >= self._submitted:
            raise _lib.FsError("BatchStylizer: nothing in flight")
        slot = self._fetched & 1
        self.ev_out[slot].synchronize()
        self._fetched += 1
        out = self.out_host[slot].numpy()
        return out.copy() if copy else out

    def stylize(self, batch):
        self.submit(batch)
        return self.fetch()
