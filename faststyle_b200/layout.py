"""Names, shapes and flat-buffer layout of the reference's variables.

The transform net's 48 variables are kept in ONE flat fp32 buffer in the
byte-sorted checkpoint key order of the reference's ``models/*.ckpt``
(SURVEY.md App. B); VGG16 conv1_1..conv4_3 follows ``libs/vgg16.py:45-173``.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

SCOPE = "img_t_net"

_conv = [("initconv_0", 9, 3, 16), ("initconv_1", 3, 16, 32), ("initconv_2", 3, 32, 64)]
_ups = [("upsample_0", 3, 64, 32), ("upsample_1", 3, 32, 16), ("upsample_2", 9, 16, 3)]


def _build():
    v = []
    for s, k, ci, co in _conv:
        v += [(s + "/INscale", (co,)), (s + "/INshift", (co,)), (s + "/W", (k, k, ci, co))]
    for r in range(5):
        s = "resblock_%d" % r
        v += [(s + "/INscale1", (64,)), (s + "/INscale2", (64,)), (s + "/INshift1", (64,)),
              (s + "/INshift2", (64,)), (s + "/W1", (3, 3, 64, 64)), (s + "/W2", (3, 3, 64, 64))]
    for s, k, ci, co in _ups:
        v += [(s + "/INscale", (co,)), (s + "/INshift", (co,)), (s + "/W", (k, k, ci, co))]
    return [(SCOPE + "/" + n, shp) for n, shp in v]


TRANSFORM_VARS = _build()                       # resize-upsampling variant (shipped ckpts)
# 'deconv' variant: conv2d_transpose weights are [k, k, cout, cin] (im_transf_net.py:173); same element
# counts, hence the same flat offsets
_DECONV_SHAPES = {SCOPE + "/upsample_0/W": (3, 3, 32, 64), SCOPE + "/upsample_1/W": (3, 3, 16, 32),
                  SCOPE + "/upsample_2/W": (9, 9, 3, 16)}
TRANSFORM_VARS_DECONV = [(n, _DECONV_SHAPES.get(n, s)) for n, s in TRANSFORM_VARS]


def transform_vars(upsample_method="resize"):
    assert upsample_method in ("resize", "deconv")
    return TRANSFORM_VARS if upsample_method == "resize" else TRANSFORM_VARS_DECONV
assert [n for n, _ in TRANSFORM_VARS] == sorted(n for n, _ in TRANSFORM_VARS)
TRANSFORM_NPARAMS = int(sum(int(np.prod(s)) for _, s in TRANSFORM_VARS))
assert TRANSFORM_NPARAMS == 424102

VGG_CONV_NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
                  "conv4_1", "conv4_2", "conv4_3"]
VGG_CHANNELS = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256),
                (256, 512), (512, 512), (512, 512)]


def transform_offsets(upsample_method="resize") -> "OrderedDict[str, tuple[int, tuple]]":
    out, off = OrderedDict(), 0
    for name, shape in transform_vars(upsample_method):
        out[name] = (off, shape)
        off += int(np.prod(shape))
    return out


def flatten_transform(params: dict, upsample_method="resize") -> np.ndarray:
    """{name: array} -> flat float32 [424102] (names with or without the scope prefix)."""
    flat = np.empty(TRANSFORM_NPARAMS, np.float32)
    for name, (off, shape) in transform_offsets(upsample_method).items():
        key = name if name in params else name[len(SCOPE) + 1:]
        if key not in params:
            raise KeyError("transform-net variable %r missing (wrong --upsample_method for this "
                           "checkpoint?)" % name)
        a = np.asarray(params[key], np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError("variable %s has shape %s, expected %s (wrong --upsample_method for "
                             "this checkpoint?)" % (name, a.shape, shape))
        flat[off:off + a.size] = a.ravel()
    return flat


def unflatten_transform(flat, upsample_method="resize") -> "OrderedDict[str, np.ndarray]":
    flat = np.asarray(flat, np.float32)
    return OrderedDict((n, flat[o:o + int(np.prod(s))].reshape(s).copy())
                       for n, (o, s) in transform_offsets(upsample_method).items())


def flatten_vgg(weights: dict) -> np.ndarray:
    """npz-style dict (conv1_1_W, conv1_1_b, ...) -> flat float32 in conv order."""
    parts = []
    for name, (cin, cout) in zip(VGG_CONV_NAMES, VGG_CHANNELS):
        w = np.asarray(weights[name + "_W"], np.float32)
        b = np.asarray(weights[name + "_b"], np.float32)
        if w.shape != (3, 3, cin, cout) or b.shape != (cout,):
            raise ValueError("VGG weight %s has shape %s/%s" % (name, w.shape, b.shape))
        parts += [w.ravel(), b.ravel()]
    return np.concatenate(parts)


def vgg_layer_index(name: str) -> int:
    """'conv3_3' or the reference's graph name 'vgg/conv3_3:0' (train.py:140-141) -> 0..9."""
    n = name
    if n.startswith("vgg/"):
        n = n[4:]
    if n.endswith(":0"):
        n = n[:-2]
    if n not in VGG_CONV_NAMES:
        raise ValueError("unsupported VGG layer %r (supported: %s)" % (name, ", ".join(VGG_CONV_NAMES)))
    return VGG_CONV_NAMES.index(n)
