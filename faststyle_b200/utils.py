"""Python-3 re-issue of the reference's ``utils.py``: OpenCV image I/O wrappers (RGB
convention) and the graph-lookup helpers ``get_layers`` / ``get_grams`` (utils.py:14-83)."""
import cv2

from . import graph
from . import ops


def imread(path):
    """cv2.imread + BGR->RGB (utils.py:14-22)."""
    img = cv2.imread(path)
    if img is None:
        raise IOError("cannot read image %r" % path)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    return img


def imresize(img, scale):
    """Cubic interpolation for up-scaling, area relation for down-scaling (utils.py:25-40)."""
    if scale > 1.0:
        img = cv2.resize(img, None, interpolation=cv2.INTER_CUBIC, fx=scale, fy=scale)
    elif scale < 1.0:
        img = cv2.resize(img, None, interpolation=cv2.INTER_AREA, fx=scale, fy=scale)
    return img


def imwrite(path, img):
    """RGB array -> cv2.imwrite (utils.py:43-52).  float32 input is saturate-rounded by OpenCV."""
    img = cv2.cvtColor(img, cv2.COLOR_RGB2BGR)
    if not cv2.imwrite(path, img):
        raise IOError("cannot write image %r" % path)


def get_layers(layer_names):
    """Get tensors by graph name, e.g. 'vgg/conv3_3:0' (utils.py:55-63)."""
    return [graph.get_tensor_by_name(name) for name in layer_names]


def get_grams(layer_names):
    """Gram matrices F^T F / (h*w*c) of the named layers (utils.py:66-83)."""
    return [ops.gram(layer) for layer in get_layers(layer_names)]
