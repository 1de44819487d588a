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
