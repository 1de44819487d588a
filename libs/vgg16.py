"""Import shim for `from libs import vgg16` (reference train.py:11); see faststyle_b200/libs/vgg16.py."""
from faststyle_b200.libs.vgg16 import vgg16  # noqa: F401
