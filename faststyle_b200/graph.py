"""Registry of named tensors, standing in for ``tf.get_default_graph().get_tensor_by_name``
(reference utils.py:55-63).  The eager vgg16 mirror registers its layer activations here under
``<scope>/<layer>:0`` (e.g. ``vgg/conv3_3:0``, train.py:140-141)."""
_tensors = {}


def clear():
    _tensors.clear()


def register(name: str, getter):
    _tensors[name] = getter


def get_tensor_by_name(name: str):
    if name not in _tensors:
        raise KeyError("The name '%s' refers to a Tensor which does not exist." % name)
    return _tensors[name]()
