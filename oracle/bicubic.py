"""TF-1.0 `ResizeBicubic` restated for the input pipeline (TEST ORACLE - see oracle/__init__.py).

Reference call site: `datapipe.py:25` - `tf.image.resize_images(image, size=resize_shape, method=2)`, i.e.
`gen_image_ops.resize_bicubic(images, size, align_corners=False)` of TensorFlow 1.0.  TensorFlow is an un-vendored
dependency (README.md:23) that cannot be installed here, so this follows the published kernel
(tensorflow/core/kernels/resize_bicubic_op.cc at r1.0) step by step, in its own arithmetic order:

* scale = float(in) / out  (no align_corners, no half-pixel centres);
* per output coordinate: in_loc = trunc(scale * out_loc) [float multiply], delta = scale * out_loc - in_loc,
  offset = lrintf(delta * 1024) (round-half-even) into a 1025-entry table of Keys cubic coefficients with
  A = -0.75, evaluated in double at x = i / 1024 and x + 1 and stored as float;
* weights = {T[offset].far, T[offset].near, T[1024 - offset].near, T[1024 - offset].far},
  indices = clamp(in_loc + {-1, 0, 1, 2}, 0, limit - 1);
* a 4x4 patch: FIRST four horizontal interpolations (one per patch row), THEN one vertical interpolation of those
  four values, each `v0*w0 + v1*w1 + v2*w2 + v3*w3` in float, left to right.

Written independently of the product's host restatement (`faststyle_b200/datapipe.py`), which interpolates rows
first; the two therefore agree to float rounding, not bit for bit.  **Parity unpinned**: no TensorFlow output
exists in the reference to check this against (tools/dump_tf1_goldens.py writes one for anyone with TF 1.x).
"""
from __future__ import annotations

import numpy as np

TABLE_SIZE = 1 << 10
A = -0.75


def _coeff_table():
    near = np.empty(TABLE_SIZE + 1, np.float32)
    far = np.empty(TABLE_SIZE + 1, np.float32)
    for i in range(TABLE_SIZE + 1):
        x = float(np.float32(i * 1.0 / TABLE_SIZE))
        near[i] = np.float32(((A + 2.0) * x - (A + 3.0)) * x * x + 1.0)
        x += 1.0
        far[i] = np.float32(((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A)
    return near, far


_NEAR, _FAR = _coeff_table()


def weights_and_indices(in_size: int, out_size: int):
    """[out,4] float32 weights and [out,4] int indices along one axis."""
    scale = np.float32(in_size) / np.float32(out_size)
    out_loc = np.arange(out_size, dtype=np.int64)
    f = (scale * out_loc.astype(np.float32)).astype(np.float32)          # float multiply
    in_loc = f.astype(np.int64)                                         # truncation (values are >= 0)
    delta = (f - in_loc.astype(np.float32)).astype(np.float32)
    offset = np.rint((delta * np.float32(TABLE_SIZE)).astype(np.float32)).astype(np.int64)   # lrintf: half to even
    w = np.stack([_FAR[offset], _NEAR[offset], _NEAR[TABLE_SIZE - offset], _FAR[TABLE_SIZE - offset]], 1)
    idx = np.stack([np.clip(in_loc + d, 0, in_size - 1) for d in (-1, 0, 1, 2)], 1)
    return w.astype(np.float32), idx


def _interp(w, v):
    """values[0]*w[0] + values[1]*w[1] + values[2]*w[2] + values[3]*w[3] in float32, left to right."""
    acc = (v[0] * w[0]).astype(np.float32)
    for k in (1, 2, 3):
        acc = (acc + (v[k] * w[k]).astype(np.float32)).astype(np.float32)
    return acc


def resize_bicubic_tf1(img, out_h: int, out_w: int) -> np.ndarray:
    """HWC uint8/float image -> HWC float32, TF-1.0 bicubic (see module docstring)."""
    img = np.asarray(img).astype(np.float32)
    H, W, _ = img.shape
    wy, iy = weights_and_indices(H, out_h)
    wx, ix = weights_and_indices(W, out_w)
    # horizontal pass on the four source rows of every output row: coeff[i][y, x, c]
    coeff = []
    for i in range(4):
        rows = img[iy[:, i]]                                            # [out_h, W, C]
        vals = [rows[:, ix[:, k]] for k in range(4)]                    # each [out_h, out_w, C]
        coeff.append(_interp([wx[None, :, k, None] for k in range(4)], vals))
    return _interp([wy[:, None, k, None] for k in range(4)], coeff)
