"""Python-3 re-issue of the reference's ``losses.py`` (content / style / TV loss) on the
B200 single-op kernels.  Inputs are NHWC / [b,C,C] tensors (torch CUDA or numpy); results
are 0-d torch CUDA tensors.  Reference: losses.py:12-97."""
import numpy as np

from . import ops


def content_loss(content_layers, target_content_layers, content_weights):
    """sum over ALL axes (incl. batch) of squared differences, x weight / (h*w*c)."""
    assert(len(target_content_layers) == len(content_layers))
    total = None
    for layer, target, w in zip(content_layers, target_content_layers, content_weights):
        _, h, wd, c = layer.shape
        term = ops.sqdiff_loss(layer, target, float(w) / float(h * wd * c))
        total = term if total is None else total + term
    return total


def style_loss(grams, target_grams, style_weights):
    """sum (gram - target)^2 x weight / (c1*c2); the [1,C,C] target broadcasts over the batch."""
    assert(len(grams) == len(target_grams))
    total = None
    for gram, target, w in zip(grams, target_grams, style_weights):
        _, c1, c2 = gram.shape
        t = np.asarray(target.detach().cpu() if hasattr(target, "detach") else target, np.float32).reshape(c1, c2)
        term = ops.style_sq_loss(gram, t, float(w) / float(c1 * c2))
        total = term if total is None else total + term
    return total


def tv_loss(X):
    """Sum of squared vertical and horizontal forward differences (no normaliser)."""
    return ops.tv_sum(X)
