"""Seeded synthetic weights.

* ``init_transform_params`` reproduces the reference's TF-1.0 initialisers
  (conv2d W ~ N(0, 0.1^2) im_transf_net.py:114; upconv2d W ~ N(0, 1) :149;
  IN scale 1 / shift 0 :233-236) with a numpy RNG (TF's own RNG stream is not
  reproducible without TF).
* ``synthetic_vgg_weights`` stands in for ``libs/vgg16_weights.npz`` (git-ignored
  in the reference and not downloadable here): He-normal conv kernels, zero
  biases, in the npz key layout ``conv1_1_W`` ... ``conv4_3_b`` (HWIO).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .layout import VGG_CONV_NAMES, VGG_CHANNELS, transform_vars


def init_transform_params(seed: int = 1, upsample_method: str = "resize") -> "OrderedDict[str, np.ndarray]":
    """'deconv': all three upsample layers are deconv2d (im_transf_net.py:57-63) whose W uses the default
    random_normal_initializer (stddev 1, :183) and the transposed [k,k,cout,cin] shape (:173)."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    unit_std = ("upsample_0", "upsample_1") if upsample_method == "resize" else ("upsample_0", "upsample_1", "upsample_2")
    for name, shape in transform_vars(upsample_method):
        leaf = name.rsplit("/", 1)[1]
        if leaf.startswith("INscale"):
            out[name] = np.ones(shape, np.float32)
        elif leaf.startswith("INshift"):
            out[name] = np.zeros(shape, np.float32)
        else:
            std = 1.0 if name.split("/")[1] in unit_std else 0.1
            out[name] = (rng.standard_normal(shape) * std).astype(np.float32)
    return out


def synthetic_vgg_weights(seed: int = 7) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, (cin, cout) in zip(VGG_CONV_NAMES, VGG_CHANNELS):
        std = np.sqrt(2.0 / (9.0 * cin))
        out[name + "_W"] = (rng.standard_normal((3, 3, cin, cout)) * std).astype(np.float32)
        out[name + "_b"] = np.zeros((cout,), np.float32)
    return out
