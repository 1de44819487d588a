"""Single-op wrappers over the C-ABI (torch CUDA tensors in, torch CUDA tensors out).
Used by the op-level mirrors in im_transf_net.py / vgg16.py / losses.py; the scripts' hot
paths use the fused composites in engine.py instead."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .engine import _require_cuda, f32, ptr, stream_ptr


def _dev(x):
    _require_cuda()
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x.to(torch.float32).contiguous()
    return f32(x, torch.device("cuda", torch.cuda.current_device()))


def _pad_c(x, mult=4):
    """Pad the channel (last) dim to a multiple of 4 with zeros (engine layout rule)."""
    c = x.shape[-1]
    p = (-c) % mult
    return (F.pad(x, (0, p)) if p else x).contiguous(), c


def same_out(n, k, s):
    return -(-n // s)


def conv2d(x, w, stride=1, padding="SAME", bias=None, relu=False):
    """tf.nn.conv2d NHWC / HWIO (im_transf_net.py:115, libs/vgg16.py:48-52)."""
    x, w = _dev(x), _dev(w)
    assert x.dim() == 4 and w.dim() == 4 and w.shape[2] == x.shape[3]
    assert padding in ("SAME", "VALID")
    N, H, W_, Ci = x.shape
    K, _, _, Co = w.shape
    xp, _ = _pad_c(x)
    wp = F.pad(w, (0, (-Co) % 4, 0, (-Ci) % 4)).contiguous()
    b = None
    if bias is not None:
        b = F.pad(_dev(bias), (0, (-Co) % 4)).contiguous()
    same = padding == "SAME"
    OH = same_out(H, K, stride) if same else (H - K) // stride + 1
    OW = same_out(W_, K, stride) if same else (W_ - K) // stride + 1
    y = torch.empty((N, OH, OW, wp.shape[3]), dtype=torch.float32, device=x.device)
    _lib.call("fs_conv2d_forward", ptr(xp), ptr(wp), ptr(b), ptr(y), N, H, W_, xp.shape[3], K, K, wp.shape[3],
              stride, 1 if same else 0, 1 if relu else 0, stream_ptr())
    return y[..., :Co].contiguous() if wp.shape[3] != Co else y


def conv2d_transpose(x, w, stride=2):
    """tf.nn.conv2d_transpose, SAME, output = input*stride; ``w`` is [k,k,cout,cin]
    (im_transf_net.py:158-190).  Computed as the data gradient of the SAME conv it transposes."""
    x, w = _dev(x), _dev(w)
    N, h, w_, Ci = x.shape
    K, _, Co, Ci2 = w.shape
    assert Ci2 == Ci
    H, W_ = h * stride, w_ * stride
    wp = F.pad(w, (0, (-Ci) % 4, 0, (-Co) % 4)).contiguous()          # conv weights [k,k,C=cout,OC=cin]
    xp, _ = _pad_c(x)
    Cp, OCp = wp.shape[2], wp.shape[3]
    out = torch.empty((N, H, W_, Cp), dtype=torch.float32, device=x.device)
    scratch = torch.empty(K * K * Cp * OCp, dtype=torch.float32, device=x.device)
    _lib.call("fs_conv2d_dgrad", ptr(xp), ptr(wp), ptr(out), ptr(scratch), N, H, W_, Cp, K, K, OCp, stride, 1,
              stream_ptr())
    return out[..., :Co].contiguous() if Cp != Co else out


def upconv2d(x, w):
    """Resize-conv: NN x4 + 3x3 stride-2 SAME conv, fused (im_transf_net.py:122-155)."""
    x, w = _dev(x), _dev(w)
    N, H, W_, Ci = x.shape
    Co = w.shape[3]
    assert w.shape[:3] == (3, 3, Ci) and Ci % 4 == 0 and Co % 4 == 0
    y = torch.empty((N, 2 * H, 2 * W_, Co), dtype=torch.float32, device=x.device)
    scratch = torch.empty(16 * Ci * Co, dtype=torch.float32, device=x.device)
    _lib.call("fs_upconv2d_forward", ptr(x), ptr(w), ptr(y), ptr(scratch), N, H, W_, Ci, Co, stream_ptr())
    return y


def inst_norm(x, scale, shift, epsilon=1e-3, act=0):
    """inst_norm (+ optional fused activation 1=relu, 2=scaled tanh) (im_transf_net.py:218-247)."""
    x = _dev(x)
    N, H, W_, C_ = x.shape
    xp, _ = _pad_c(x)
    Cp = xp.shape[3]
    g = F.pad(_dev(scale), (0, Cp - C_)).contiguous()
    b = F.pad(_dev(shift), (0, Cp - C_)).contiguous()
    y = torch.empty_like(xp)
    stats = torch.empty(2 * N * Cp, dtype=torch.float32, device=x.device)
    scratch = torch.empty(N * 64 * 2 * Cp, dtype=torch.float64, device=x.device)
    _lib.call("fs_instnorm_forward", ptr(xp), ptr(g), ptr(b), ptr(y), ptr(stats), ptr(scratch), N, H, W_, Cp,
              C.c_float(epsilon), act, stream_ptr())
    return y[..., :C_].contiguous() if Cp != C_ else y


def max_pool(x):
    x = _dev(x)
    N, H, W_, C_ = x.shape
    assert C_ % 4 == 0
    y = torch.empty((N, (H + 1) // 2, (W_ + 1) // 2, C_), dtype=torch.float32, device=x.device)
    _lib.call("fs_maxpool2x2", ptr(x), ptr(y), N, H, W_, C_, stream_ptr())
    return y


def gram(f):
    """G = F^T F / (h*w*c) per sample (utils.py:76-81)."""
    f = _dev(f)
    N, H, W_, C_ = f.shape
    assert C_ % 4 == 0
    n = _lib.load().fs_gram_scratch_floats(N, C_)
    scratch = torch.empty(n, dtype=torch.float32, device=f.device)
    g = torch.empty((N, C_, C_), dtype=torch.float32, device=f.device)
    _lib.call("fs_gram_forward", ptr(f), ptr(g), ptr(scratch), C.c_longlong(n), N, H, W_, C_, stream_ptr())
    return g


def _loss_bufs(dev):
    return torch.empty(4, dtype=torch.float64, device=dev), torch.empty(4, dtype=torch.float32, device=dev)


def sqdiff_loss(a, b, scale):
    a, b = _dev(a), _dev(b)
    assert a.shape == b.shape
    n = a.numel()
    if n % 4:
        a = F.pad(a.flatten(), (0, (-n) % 4)); b = F.pad(b.flatten(), (0, (-n) % 4))
    acc, out = _loss_bufs(a.device)
    _lib.call("fs_loss_sqdiff", ptr(a.contiguous()), ptr(b.contiguous()), C.c_longlong(a.numel()),
              C.c_double(scale), ptr(acc), ptr(out), stream_ptr())
    return out[0]


def style_sq_loss(G, T, scale):
    G, T = _dev(G), _dev(T)
    N, c1, c2 = G.shape
    acc, out = _loss_bufs(G.device)
    _lib.call("fs_loss_style", ptr(G), ptr(T.reshape(-1).contiguous()), N, c1 * c2, C.c_double(scale), ptr(acc),
              ptr(out), stream_ptr())
    return out[1]


def tv_sum(Y):
    Y = _dev(Y)
    N, H, W_, C_ = Y.shape
    assert C_ == 3, "tv_loss is defined on RGB tensors (losses.py:70-97)"
    acc, out = _loss_bufs(Y.device)
    _lib.call("fs_loss_tv", ptr(Y), N, H, W_, ptr(acc), ptr(out), stream_ptr())
    return out[2]


def resize_bicubic_tf1(img_u8, out_h, out_w, out=None):
    """tf.image.resize_images(img, [out_h, out_w], method=2) of TF 1.0 on the GPU (reference datapipe.py:25):
    ``img_u8`` is one decoded HWC uint8 RGB image (numpy or CUDA tensor); returns / fills a float32 CUDA
    tensor [out_h, out_w, 3]."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(img_u8, np.ndarray):
        img_u8 = torch.from_numpy(np.ascontiguousarray(img_u8))
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 3 or img_u8.shape[2] != 3:
        raise _lib.FsError("resize_bicubic_tf1 expects an HWC uint8 RGB image")
    src = img_u8.to(dev, non_blocking=True).contiguous()
    if out is None:
        out = torch.empty((out_h, out_w, 3), dtype=torch.float32, device=dev)
    _lib.call("fs_resize_bicubic_tf1_u8", ptr(src), ptr(out), int(src.shape[0]), int(src.shape[1]), int(out_h),
              int(out_w), stream_ptr())
    return out
