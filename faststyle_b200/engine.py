"""Python handle over the C-ABI engine (torch tensors are device storage only).

``Engine`` owns a plan + one workspace tensor; all arithmetic happens inside
libfaststyle_b200.so on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ENG_DECONV, ENG_TRANSFORM, ENG_TRANSFORM_BWD, ENG_VGG, ENG_VGG_BWD, FsError, LossConfigC
from .layout import (TRANSFORM_NPARAMS, VGG_CONV_NAMES, flatten_transform, flatten_vgg,
                     vgg_layer_index)


def _require_cuda():
    if not torch.cuda.is_available():
        raise FsError("faststyle_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        raise FsError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(x, device) -> torch.Tensor:
    """numpy / torch (any dtype incl. uint8) -> contiguous float32 CUDA tensor."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def make_loss_config(content_layers: Sequence[str], content_weights: Sequence[float],
                     style_layers: Sequence[str], style_weights: Sequence[float], beta: float):
    if len(content_layers) != len(content_weights):
        raise AssertionError("len(content_weights) != len(content_layers)")      # losses.py:23
    if len(style_layers) != len(style_weights):
        raise AssertionError("len(style_weights) != len(style_layers)")          # losses.py:53
    cfg = LossConfigC()
    cfg.n_content = len(content_layers)
    for i, (n, w) in enumerate(zip(content_layers, content_weights)):
        cfg.content_layer[i] = vgg_layer_index(n)
        cfg.content_w[i] = float(w)
    cfg.n_style = len(style_layers)
    for i, (n, w) in enumerate(zip(style_layers, style_weights)):
        cfg.style_layer[i] = vgg_layer_index(n)
        cfg.style_w[i] = float(w)
    cfg.beta = float(beta)
    return cfg


def _mask(layers) -> int:
    m = 0
    for n in layers:
        m |= 1 << vgg_layer_index(n)
    return m


def pack_vgg(weights, device) -> torch.Tensor:
    """npz-style VGG weights -> engine-packed device buffer (libs/vgg16.py:257-266)."""
    _require_cuda()
    lib = _lib.load()
    flat = torch.from_numpy(flatten_vgg(weights)).to(device)
    packed = torch.empty(lib.fs_vgg_packed_floats(), dtype=torch.float32, device=device)
    with torch.cuda.device(flat.device):
        _lib.call("fs_vgg_pack", ptr(flat), ptr(packed), stream_ptr())
        torch.cuda.current_stream().synchronize()
    return packed


class Engine:
    """Plan + workspace for batch ``N`` of ``H x W`` RGB images."""

    def __init__(self, N, H, W, transform=False, transform_bwd=False, vgg=False, vgg_bwd=False,
                 content_layers=(), style_layers=(), device="cuda:0", deconv=False):
        _require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.N, self.H, self.W = int(N), int(H), int(W)
        flags = (ENG_TRANSFORM if transform or transform_bwd else 0) | (ENG_TRANSFORM_BWD if transform_bwd else 0) \
            | (ENG_VGG if vgg or vgg_bwd else 0) | (ENG_VGG_BWD if vgg_bwd else 0) | (ENG_DECONV if deconv else 0)
        self.flags = flags
        self._h = C.c_void_p()
        _lib.call("fs_engine_create", self.N, self.H, self.W, flags, _mask(content_layers),
                  _mask(style_layers), C.byref(self._h))
        nbytes = self.lib.fs_engine_workspace_bytes(self._h)
        self.workspace_bytes = int(nbytes)
        self.workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        _lib.call("fs_engine_bind", self._h, ptr(self.workspace), C.c_size_t(self.workspace.numel()))
        oh, ow = C.c_int(), C.c_int()
        _lib.call("fs_engine_output_dims", self._h, C.byref(oh), C.byref(ow))
        self.OH, self.OW = oh.value, ow.value

    def set_tensor_path(self, enabled: bool):
        """True (default): tcgen05 split-bf16 path for the 64-channel-multiple 3x3 convs;
        False: exact-fp32 FFMA path everywhere."""
        _lib.call("fs_engine_set_tensor_path", self._h, 1 if enabled else 0)

    def keep_activations(self, keep: bool = True):
        """Also write the fp32 activations whose only consumers read split-bf16 planes (skipped by default); call
        before a forward pass whose intermediate activations are inspected with ``transform_activation``."""
        _lib.call("fs_engine_keep_activations", self._h, 1 if keep else 0)

    def set_frozen_weights(self, frozen: bool):
        """Inference with fixed parameters: prepare the transform weights once (see fs_engine_set_frozen_weights)."""
        _lib.call("fs_engine_set_frozen_weights", self._h, 1 if frozen else 0)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.fs_engine_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # -------------------------------------------------------------- helpers
    def _x3(self, x, H=None, W=None):
        x = f32(x, self.device)
        exp = (self.N, H or self.H, W or self.W, 3)
        if tuple(x.shape) != exp:
            raise FsError("expected an NHWC tensor of shape %s, got %s" % (exp, tuple(x.shape)))
        return x

    def _tg_array(self, target_grams):
        arr = (C.c_void_p * max(len(target_grams), 1))()
        for i, t in enumerate(target_grams):
            arr[i] = t.data_ptr()
        return arr

    # -------------------------------------------------------------- composites
    def transform_forward(self, params_flat: torch.Tensor, x) -> torch.Tensor:
        x = self._x3(x)
        y = torch.empty((self.N, self.OH, self.OW, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("fs_transform_forward", self._h, ptr(params_flat), ptr(x), ptr(y), stream_ptr())
        return y

    def vgg_forward(self, packed, img, upto="conv4_3"):
        img = self._x3(img, self.OH, self.OW)
        with torch.cuda.device(self.device):
            _lib.call("fs_vgg_forward", self._h, ptr(packed), ptr(img), vgg_layer_index(upto), stream_ptr())

    def vgg_activation(self, name) -> torch.Tensor:
        """Copy of a saved VGG activation as an NHWC tensor."""
        p, h, w, c = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        _lib.call("fs_engine_vgg_activation", self._h, vgg_layer_index(name), C.byref(p), C.byref(h),
                  C.byref(w), C.byref(c))
        return self._view(p.value, (self.N, h.value, w.value, c.value)).clone()

    def transform_activation(self, conv_index, stage) -> torch.Tensor:
        p, h, w, c = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        _lib.call("fs_engine_transform_activation", self._h, conv_index, stage, C.byref(p), C.byref(h),
                  C.byref(w), C.byref(c))
        return self._view(p.value, (self.N, h.value, w.value, c.value)).clone()

    def _view(self, addr, shape):
        off = addr - self.workspace.data_ptr()
        n = int(np.prod(shape))
        return self.workspace[off:off + 4 * n].view(torch.float32).view(shape)

    def vgg_grams(self, packed, img, layers):
        img = self._x3(img, self.OH, self.OW)
        idx = [vgg_layer_index(n) for n in layers]
        from .layout import VGG_CHANNELS
        outs = [torch.empty((self.N, VGG_CHANNELS[i][1], VGG_CHANNELS[i][1]), dtype=torch.float32,
                            device=self.device) for i in idx]
        with torch.cuda.device(self.device):
            _lib.call("fs_vgg_grams", self._h, ptr(packed), ptr(img), len(idx), (C.c_int * len(idx))(*idx),
                      self._tg_array(outs), stream_ptr())
        return outs

    def set_content_targets(self, packed, img, cfg):
        img = self._x3(img, self.OH, self.OW)
        with torch.cuda.device(self.device):
            _lib.call("fs_vgg_set_content_targets", self._h, ptr(packed), ptr(img), C.byref(cfg), stream_ptr())

    def perceptual_loss(self, packed, img, cfg, target_grams, need_grad=True):
        img = self._x3(img, self.OH, self.OW)
        losses = torch.empty(4, dtype=torch.float32, device=self.device)
        grad = torch.empty_like(img) if need_grad else None
        with torch.cuda.device(self.device):
            _lib.call("fs_perceptual_loss", self._h, ptr(packed), ptr(img), C.byref(cfg),
                      self._tg_array(target_grams), ptr(losses), ptr(grad), stream_ptr())
        return losses, grad

    def train_fwd_bwd(self, params_flat, packed, x, cfg, target_grams, grads=None, losses=None, y=None):
        x = self._x3(x)
        if grads is None:
            grads = torch.empty(TRANSFORM_NPARAMS, dtype=torch.float32, device=self.device)
        if losses is None:
            losses = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("fs_train_fwd_bwd", self._h, ptr(params_flat), ptr(packed), ptr(x), C.byref(cfg),
                      self._tg_array(target_grams), ptr(grads), ptr(losses), ptr(y), stream_ptr())
        return grads, losses


class TFAdam:
    """Fused tf.train.AdamOptimizer on a flat device buffer (train.py:203)."""

    def __init__(self, params_flat: torch.Tensor, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        _require_cuda()
        self.p = params_flat
        self.m = torch.zeros_like(params_flat)
        self.v = torch.zeros_like(params_flat)
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=params_flat.device)
        self.lr, self.b1, self.b2, self.eps = float(lr), float(beta1), float(beta2), float(eps)

    def step(self, grads: torch.Tensor):
        with torch.cuda.device(self.p.device):
            _lib.call("fs_adam_step", ptr(self.p), ptr(grads), ptr(self.m), ptr(self.v),
                      C.c_longlong(self.p.numel()), self.lr, self.b1, self.b2, self.eps,
                      ptr(self.step_counter), stream_ptr())


def params_to_device(params: dict, device, upsample_method="resize") -> torch.Tensor:
    return torch.from_numpy(flatten_transform(params, upsample_method)).to(device)
