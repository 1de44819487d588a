"""Torch-CPU restatement of the reference graph (TEST ORACLE — see oracle/__init__.py).

Activations are NHWC at the API (as in the reference) and NCHW internally.
Weights are HWIO ``[kh, kw, cin, cout]`` exactly as stored in the reference's
checkpoints / npz.  Every function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

VGG_MEAN = (123.68, 116.779, 103.939)          # libs/vgg16.py:41
VGG_LAYERS = [                                  # libs/vgg16.py:45-173 (conv1_1..conv4_3)
    ("conv1_1", 3, 64), ("conv1_2", 64, 64), ("pool1",),
    ("conv2_1", 64, 128), ("conv2_2", 128, 128), ("pool2",),
    ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("pool3",),
    ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
]


# ---------------------------------------------------------------- TF semantics
def same_pad(n: int, k: int, s: int):
    """TF 'SAME' padding: out=ceil(n/s); total=max((out-1)s+k-n,0); before=total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _t(x, dtype):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(dtype)


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def nchw_to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def conv2d_tf(x, w_hwio, stride: int, padding: str):
    """tf.nn.conv2d on NCHW ``x`` with HWIO weights (im_transf_net.py:115-118)."""
    k = w_hwio.shape[0]
    w = w_hwio.permute(3, 2, 0, 1)
    if padding == "SAME":
        pt, pb = same_pad(x.shape[2], k, stride)
        pl, pr = same_pad(x.shape[3], k, stride)
        x = F.pad(x, (pl, pr, pt, pb))
    elif padding != "VALID":
        raise ValueError(padding)
    return F.conv2d(x, w, stride=stride)


def resize_nearest_tf1(x, new_h: int, new_w: int):
    """tf.image.resize_images(method=1) TF-1.0 legacy mapping src=floor(dst*in/out)
    (im_transf_net.py:140-142)."""
    h, w = x.shape[2], x.shape[3]
    iy = torch.div(torch.arange(new_h) * h, new_h, rounding_mode="floor").long()
    ix = torch.div(torch.arange(new_w) * w, new_w, rounding_mode="floor").long()
    return x[:, :, iy][:, :, :, ix]


def max_pool_same(x):
    """tf.nn.max_pool k2 s2 SAME (libs/vgg16.py:67-71); odd sizes pad bottom/right."""
    ph, pw = x.shape[2] % 2, x.shape[3] % 2
    if ph or pw:
        x = F.pad(x, (0, pw, 0, ph), value=float("-inf"))
    return F.max_pool2d(x, 2, 2)


# ---------------------------------------------------------------- transform net
def reflect_pad(x, padsize: int):
    """im_transf_net.py:78-88 (tf.pad REFLECT == torch 'reflect')."""
    return F.pad(x, (padsize,) * 4, mode="reflect")


def inst_norm(x, scale, shift, epsilon=1e-3):
    """im_transf_net.py:218-247: biased moments over H,W; eps inside sqrt."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    normed = (x - mean) / torch.sqrt(var + epsilon)
    return scale.view(1, -1, 1, 1) * normed + shift.view(1, -1, 1, 1)


def scaled_tanh(x):
    """im_transf_net.py:202-215."""
    return (255.0 * torch.tanh(x) + 255.0) / 2.0


def upconv2d(x, w, stride: int):
    """im_transf_net.py:122-155: NN-resize by stride**2 then conv stride SAME."""
    up = resize_nearest_tf1(x, x.shape[2] * stride ** 2, x.shape[3] * stride ** 2)
    return conv2d_tf(up, w, stride, "SAME")


def deconv2d(x, w_hwoi, stride: int):
    """im_transf_net.py:158-190: tf.nn.conv2d_transpose k3 SAME; W is [k,k,cout,cin].
    Defined as the gradient of conv2d(SAME) w.r.t. its input of size in*stride."""
    k = w_hwoi.shape[0]
    oh, ow = x.shape[2] * stride, x.shape[3] * stride
    pt, _ = same_pad(oh, k, stride)
    pl, _ = same_pad(ow, k, stride)
    # conv_transpose2d weight layout: [cin(of transpose)=x channels, cout, kh, kw]
    wt = w_hwoi.permute(3, 2, 0, 1)
    full = F.conv_transpose2d(x, wt, stride=stride)
    need_h, need_w = pt + oh, pl + ow
    if full.shape[2] < need_h or full.shape[3] < need_w:
        full = F.pad(full, (0, max(need_w - full.shape[3], 0), 0, max(need_h - full.shape[2], 0)))
    return full[:, :, pt:pt + oh, pl:pl + ow]


def res_layer(x, p, scope):
    """im_transf_net.py:250-276."""
    h = conv2d_tf(x, p[scope + "/W1"], 1, "VALID")
    h = F.relu(inst_norm(h, p[scope + "/INscale1"], p[scope + "/INshift1"]))
    h = conv2d_tf(h, p[scope + "/W2"], 1, "VALID")
    h = inst_norm(h, p[scope + "/INscale2"], p[scope + "/INshift2"])
    return h + x[:, :, 2:-2, 2:-2]


def _params(params, dtype, prefix="img_t_net/"):
    out = {}
    for k, v in params.items():
        kk = k[len(prefix):] if k.startswith(prefix) else k
        out[kk] = v if isinstance(v, torch.Tensor) and v.dtype == dtype else _t(v, dtype)
    return out


def create_net(X, params, upsample_method="resize", dtype=torch.float32, taps=None):
    """im_transf_net.py:14-75.  ``X`` NHWC 0..255; returns NHWC 0..255.
    ``taps`` (optional dict) receives NHWC copies of intermediate activations."""
    assert upsample_method in ("deconv", "resize")
    p = _params(params, dtype)
    x = nhwc_to_nchw(_t(X, dtype))

    def tap(name, t):
        if taps is not None:
            taps[name] = nchw_to_nhwc(t).detach()
        return t

    h = reflect_pad(x, 40)
    for i, (s, pad) in enumerate([(1, "SAME"), (2, "SAME"), (2, "SAME")]):
        sc = "initconv_%d" % i
        h = conv2d_tf(h, p[sc + "/W"], s, pad)
        tap(sc + "/conv", h)
        h = tap(sc, F.relu(inst_norm(h, p[sc + "/INscale"], p[sc + "/INshift"])))
    for i in range(5):
        h = tap("resblock_%d" % i, res_layer(h, p, "resblock_%d" % i))
    up = upconv2d if upsample_method == "resize" else deconv2d
    for i in range(2):
        sc = "upsample_%d" % i
        h = up(h, p[sc + "/W"], 2)
        tap(sc + "/conv", h)
        h = tap(sc, F.relu(inst_norm(h, p[sc + "/INscale"], p[sc + "/INshift"])))
    sc = "upsample_2"
    if upsample_method == "resize":
        h = conv2d_tf(h, p[sc + "/W"], 1, "SAME")
    else:
        h = deconv2d(h, p[sc + "/W"], 1)
    tap(sc + "/conv", h)
    h = scaled_tanh(inst_norm(h, p[sc + "/INscale"], p[sc + "/INshift"]))
    return nchw_to_nhwc(h)


# ---------------------------------------------------------------- VGG16
def vgg16_layers(imgs_nhwc, weights, upto="conv4_3", dtype=torch.float32, nchw_in=False):
    """libs/vgg16.py:36-173: mean-subtract, 3x3 SAME conv + bias + ReLU, 2x2 SAME
    max-pool.  ``weights``: npz-style dict ``conv1_1_W`` [3,3,cin,cout], ``conv1_1_b``.
    Returns OrderedDict name -> NCHW activation (post-ReLU), built up to ``upto``."""
    x = imgs_nhwc if nchw_in else nhwc_to_nchw(_t(imgs_nhwc, dtype))
    mean = torch.tensor(VGG_MEAN, dtype=dtype).view(1, 3, 1, 1)
    h = x - mean
    out = OrderedDict()
    for spec in VGG_LAYERS:
        name = spec[0]
        if name.startswith("pool"):
            h = max_pool_same(h)
        else:
            w = _t(weights[name + "_W"], dtype)
            b = _t(weights[name + "_b"], dtype)
            h = F.relu(conv2d_tf(h, w, 1, "SAME") + b.view(1, -1, 1, 1))
        out[name] = h
        if name == upto:
            break
    return out


def gram(layer_nchw):
    """utils.py:66-83: F=[b,hw,c]; G = F^T F / (h*w*c)."""
    b, c, h, w = layer_nchw.shape
    f = layer_nchw.reshape(b, c, h * w)
    return torch.bmm(f, f.transpose(1, 2)) / float(h * w * c)


def content_loss(layers, targets, weights):
    """losses.py:12-40 (sum over ALL axes incl. batch; normaliser h*w*c)."""
    assert len(layers) == len(targets)
    total = 0.0
    for l, t, w in zip(layers, targets, weights):
        _, c, h, ww = l.shape
        total = total + w * torch.sum((l - t) ** 2) / float(h * ww * c)
    return total


def style_loss(grams, target_grams, weights):
    """losses.py:43-67 (target [1,C,C] broadcast over the batch; normaliser c*c)."""
    assert len(grams) == len(target_grams)
    total = 0.0
    for g, t, w in zip(grams, target_grams, weights):
        size = g.shape[1] * g.shape[2]
        total = total + w * torch.sum((g - t) ** 2) / float(size)
    return total


def tv_loss(x_nchw):
    """losses.py:70-97: sum of squared forward differences, no normaliser."""
    vdiff = x_nchw[:, :, :-1, :] - x_nchw[:, :, 1:, :]
    hdiff = x_nchw[:, :, :, :-1] - x_nchw[:, :, :, 1:]
    return torch.sum(hdiff ** 2) + torch.sum(vdiff ** 2)


def style_target_grams(style_img_nhwc, vgg_weights, style_layers, dtype=torch.float32):
    """train.py:143-151: Grams of the style image, each [1,C,C]."""
    with torch.no_grad():
        L = vgg16_layers(style_img_nhwc, vgg_weights, "conv4_3", dtype)
        return [gram(L[n]) for n in style_layers]


# ---------------------------------------------------------------- TF Adam
class TFAdam:
    """tf.train.AdamOptimizer defaults (train.py:203): epsilon is added to sqrt(v)
    WITHOUT bias correction; lr_t = lr*sqrt(1-b2^t)/(1-b1^t)."""

    def __init__(self, params: dict, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params: dict, grads: dict):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for k in params:
                g = grads[k]
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                params[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


# ---------------------------------------------------------------- train step
def train_losses(X_nhwc, params, vgg_weights, target_grams, content_layers=("conv3_3",),
                 style_layers=("conv1_2", "conv2_2", "conv3_3", "conv4_3"),
                 content_weights=(1.0,), style_weights=(5.0, 5.0, 5.0, 5.0), beta=0.0,
                 upsample_method="resize", dtype=torch.float32):
    """train.py:158-184 + 250-255: returns (loss, content, style, beta*tv, Y_nhwc).
    Content targets are VGG(content layer) of the RAW batch (train.py:250-251)."""
    with torch.no_grad():
        tgt = vgg16_layers(X_nhwc, vgg_weights, "conv4_3", dtype)
        content_targets = [tgt[n] for n in content_layers]
    Y = create_net(X_nhwc, params, upsample_method, dtype)
    Yc = nhwc_to_nchw(Y)
    L = vgg16_layers(Yc, vgg_weights, "conv4_3", dtype, nchw_in=True)
    c = content_loss([L[n] for n in content_layers], content_targets, content_weights)
    s = style_loss([gram(L[n]) for n in style_layers],
                   [_t(t, dtype) for t in target_grams], style_weights)
    tv = tv_loss(Yc)
    loss = c + s + beta * tv
    return loss, c, s, beta * tv, Y


def train_grads(X_nhwc, params, vgg_weights, target_grams, dtype=torch.float32, **kw):
    """Loss scalars + gradients of the 48 transform-net variables (the reference
    delegates this to tf.gradients via minimize, train.py:198-204)."""
    p = {k: _t(v, dtype).clone().requires_grad_(True) for k, v in params.items()}
    loss, c, s, tv, Y = train_losses(X_nhwc, p, vgg_weights, target_grams, dtype=dtype, **kw)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items()}
    return dict(loss=loss.detach(), content=torch.as_tensor(c).detach(),
                style=torch.as_tensor(s).detach(), tv=torch.as_tensor(tv).detach(),
                Y=Y.detach(), grads=grads)


def slow_style_grads(X_var_nhwc, content_targets, vgg_weights, target_grams,
                     content_layers=("conv3_3",),
                     style_layers=("conv1_2", "conv2_2", "conv3_3", "conv4_3"),
                     content_weights=(1.0,), style_weights=(5.0, 5.0, 5.0, 5.0), beta=1e-4,
                     dtype=torch.float32):
    """slow_style.py:116-154: loss and pixel gradient of the variable image."""
    x = nhwc_to_nchw(_t(X_var_nhwc, dtype)).clone().requires_grad_(True)
    L = vgg16_layers(x, vgg_weights, "conv4_3", dtype, nchw_in=True)
    c = content_loss([L[n] for n in content_layers], [_t(t, dtype) for t in content_targets],
                     content_weights)
    s = style_loss([gram(L[n]) for n in style_layers], [_t(t, dtype) for t in target_grams],
                   style_weights)
    tv = tv_loss(x)
    loss = c + s + beta * tv
    loss.backward()
    return dict(loss=loss.detach(), content=c.detach(), style=s.detach(), tv=(beta * tv).detach(),
                grad=nchw_to_nhwc(x.grad.detach()))
