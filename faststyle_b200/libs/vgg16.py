"""Python-3 re-issue of the reference's ``libs/vgg16.py`` class on the B200 engine.

Eager instead of graph-mode: the constructor takes the NHWC image batch, the layer
attributes ``conv1_1`` ... ``conv4_3`` (post-ReLU, NHWC torch CUDA tensors) are evaluated on
first access with the current ``parameters``.  conv5_x / pool4 / pool5 of the reference
(libs/vgg16.py:176-220) are never fetched by any script and are not built.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import graph
from .. import variables as V
from ..engine import Engine, f32, pack_vgg
from ..layout import VGG_CHANNELS, VGG_CONV_NAMES


class vgg16:
    def __init__(self, imgs, weights=None, sess=None):
        self.imgs = imgs
        self.scope = V.current_scope()
        self.convlayers()
        if weights is not None:
            self.load_weights(weights, sess)

    def convlayers(self):
        # truncated_normal(stddev=0.1) kernels / zero biases until load_weights (libs/vgg16.py:46-52)
        rng = np.random.RandomState(0)
        self.parameters = []
        for cin, cout in VGG_CHANNELS:
            w = np.clip(rng.standard_normal((3, 3, cin, cout)), -2, 2).astype(np.float32) * 0.1
            self.parameters += [w, np.zeros((cout,), np.float32)]
        self._packed = None
        self._engine = None
        self._valid = False
        for name in VGG_CONV_NAMES:
            graph.register((self.scope + "/" if self.scope else "") + name + ":0",
                           (lambda n: (lambda: getattr(self, n)))(name))

    def load_weights(self, weight_file, sess=None):
        """np.load the npz, zip its SORTED keys onto ``parameters`` and stop at the first key
        containing 'fc' (libs/vgg16.py:257-266); conv5_x entries are skipped (not built here)."""
        weights = np.load(weight_file) if isinstance(weight_file, str) else weight_file
        keys = sorted(weights.keys())
        i = 0
        for k in keys:
            if 'fc' in k:
                break
            if i < len(self.parameters):
                v = np.asarray(weights[k], np.float32)
                if v.shape != self.parameters[i].shape:
                    raise ValueError("VGG weight %s has shape %s, expected %s" % (k, v.shape, self.parameters[i].shape))
                self.parameters[i] = v
            i += 1
        if i < len(self.parameters):
            raise ValueError("weight file holds only %d of the %d conv1_1..conv4_3 tensors" % (i, len(self.parameters)))
        self._packed = None
        self._valid = False

    def set_input(self, imgs):
        self.imgs = imgs
        self._valid = False

    def _weights_dict(self):
        d = {}
        for i, name in enumerate(VGG_CONV_NAMES):
            d[name + "_W"] = self.parameters[2 * i]
            d[name + "_b"] = self.parameters[2 * i + 1]
        return d

    def _run(self):
        dev = torch.device("cuda", torch.cuda.current_device())
        x = f32(self.imgs, dev)
        N, H, W_, _ = x.shape
        if self._engine is None or (self._engine.N, self._engine.H, self._engine.W) != (N, H, W_):
            self._engine = Engine(N, H, W_, vgg=True, style_layers=VGG_CONV_NAMES, device=dev)
        if self._packed is None:
            self._packed = pack_vgg(self._weights_dict(), dev)
        self._engine.vgg_forward(self._packed, x, "conv4_3")
        self._valid = True

    def __getattr__(self, name):
        if name in VGG_CONV_NAMES:
            if not self.__dict__.get("_valid", False):
                self._run()
            return self._engine.vgg_activation(name)
        raise AttributeError(name)

    def grams(self, layer_names):
        """Gram matrices of the named layers through the fused C-ABI entry (utils.get_grams)."""
        if self._packed is None:
            self._packed = pack_vgg(self._weights_dict(), torch.device("cuda", torch.cuda.current_device()))
        dev = self._packed.device
        x = f32(self.imgs, dev)
        N, H, W_, _ = x.shape
        if self._engine is None or (self._engine.N, self._engine.H, self._engine.W) != (N, H, W_):
            self._engine = Engine(N, H, W_, vgg=True, style_layers=VGG_CONV_NAMES, device=dev)
        self._valid = False
        return self._engine.vgg_grams(self._packed, x, layer_names)
