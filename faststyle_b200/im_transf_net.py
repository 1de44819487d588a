"""Python-3 re-issue of the reference's ``im_transf_net.py`` public functions, running on
the B200 engine.  Same names, arguments and defaults; tensors are NHWC float32 (numpy or
torch) in, torch CUDA NHWC float32 out.  Variables come from ``variables`` (TF-style
scopes); missing variables are created with the reference's initialisers.

``create_net`` runs the whole network through the fused C-ABI composite
(``fs_transform_forward``); the per-op functions call the single-op entry points and exist
for op-level use and tests.  Reference: im_transf_net.py:14-276.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from . import variables as V
from .engine import Engine, f32, params_to_device
from .layout import SCOPE, transform_vars

_RNG = np.random.RandomState(0)
_engines = {}


def _normal(std):
    return lambda shape: (_RNG.standard_normal(shape) * std).astype(np.float32)


def _ones(shape):
    return np.ones(shape, np.float32)


def _zeros(shape):
    return np.zeros(shape, np.float32)


def create_net(X, upsample_method='deconv'):
    """Creates (evaluates) the transformation network on the NHWC batch ``X`` (values 0..255).

    :param X  NxHxWx3 array / tensor
    :param upsample_method  'deconv' or 'resize'.  The shipped checkpoints and both CLIs default to
        'resize' (fused resize-convolution).  'deconv' (conv2d_transpose, im_transf_net.py:57-63,158-190)
        is supported for inference here and for training through train.py --upsample_method deconv.
    """
    assert(upsample_method in ['deconv', 'resize'])
    scope = V.current_scope()
    if scope != SCOPE:
        raise ValueError("create_net must be called inside variable_scope('%s') so that variable names match "
                         "the reference's checkpoints (got scope %r)" % (SCOPE, scope))
    # variables are declared relative to the current scope (created with the reference's
    # initialisers when a checkpoint has not been restored)
    params = {}
    for full, shape in transform_vars(upsample_method):
        rel = full[len(SCOPE) + 1:]
        params[full] = V.get_variable(rel, shape, _initializer_for(rel, upsample_method))
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    x = f32(X, dev) if dev is not None else X
    N, H, W_, C_ = x.shape
    assert C_ == 3, "create_net expects RGB input (N x H x W x 3)"
    key = (N, H, W_, str(dev), upsample_method)
    eng = _engines.get(key)
    if eng is None:
        _engines.clear()                      # keep at most one cached plan
        eng = _engines[key] = Engine(N, H, W_, transform=True, device=dev, deconv=upsample_method == 'deconv')
    return eng.transform_forward(params_to_device(params, dev, upsample_method), x)


def _initializer_for(rel, upsample_method='resize'):
    leaf = rel.rsplit("/", 1)[1]
    if leaf.startswith("INscale"):
        return _ones
    if leaf.startswith("INshift"):
        return _zeros
    # upconv2d / deconv2d use the default random_normal_initializer (stddev 1, im_transf_net.py:149,183);
    # in the 'deconv' variant upsample_2 is a deconv2d too (:63)
    unit = ("upsample_0", "upsample_1") + (("upsample_2",) if upsample_method == 'deconv' else ())
    return _normal(1.0 if rel.split("/")[0] in unit else 0.1)


# --------------------------------------------------------------------------- per-op mirrors
def reflect_pad(X, padsize):
    """Pre-net padding (tf.pad REFLECT, im_transf_net.py:78-88).  Index plumbing only."""
    x = ops._dev(X)
    return torch.nn.functional.pad(x.permute(0, 3, 1, 2), (padsize,) * 4, mode="reflect").permute(0, 2, 3, 1).contiguous()


def conv2d(X, n_ch_in, n_ch_out, kernel_size, strides, name=None, padding='SAME'):
    """Convolutional layer without bias (im_transf_net.py:91-119)."""
    if name is None:
        name = 'W'
    W = V.get_variable(name, [kernel_size, kernel_size, n_ch_in, n_ch_out], _normal(0.1))
    assert strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    return ops.conv2d(X, W, strides[1], padding)


def upconv2d(X, n_ch_in, n_ch_out, kernel_size, strides):
    """Resize (nearest, x strides**2) then convolve with stride `strides` (im_transf_net.py:122-155)."""
    assert kernel_size == 3 and list(strides) == [1, 2, 2, 1], "the fused resize-conv kernel covers k=3, stride 2"
    W = V.get_variable('W', [kernel_size, kernel_size, n_ch_in, n_ch_out], _normal(1.0))
    return ops.upconv2d(X, W)


def deconv2d(X, n_ch_in, n_ch_out, kernel_size, strides):
    """Transposed convolution, SAME, output = input * stride (im_transf_net.py:158-190).
    Note the reversed channel order of the weight: [k, k, n_ch_out, n_ch_in]."""
    assert strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    W = V.get_variable('W', [kernel_size, kernel_size, n_ch_out, n_ch_in], _normal(1.0))
    return ops.conv2d_transpose(X, W, strides[1])


def relu(X):
    return torch.relu(ops._dev(X))


def scaled_tanh(X):
    """(255*tanh(x)+255)/2 (im_transf_net.py:202-215)."""
    return (255.0 * torch.tanh(ops._dev(X)) + 255.0) / 2.0


def inst_norm(inputs, epsilon=1e-3, suffix=''):
    """Instance normalisation with trainable scale/shift (im_transf_net.py:218-247)."""
    c = inputs.shape[3]
    scale = V.get_variable('INscale' + suffix, [c], _ones)
    shift = V.get_variable('INshift' + suffix, [c], _zeros)
    return ops.inst_norm(inputs, scale, shift, epsilon)


def res_layer(X, n_ch, kernel_size, strides):
    """Residual block (im_transf_net.py:250-276)."""
    h = conv2d(X, n_ch, n_ch, kernel_size, strides, name='W1', padding='VALID')
    h = relu(inst_norm(h, suffix='1'))
    h = conv2d(h, n_ch, n_ch, kernel_size, strides, name='W2', padding='VALID')
    h = inst_norm(h, suffix='2')
    X = ops._dev(X)
    return h + X[:, 2:-2, 2:-2, :]
