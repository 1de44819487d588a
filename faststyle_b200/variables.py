"""Minimal stand-in for the TF-1 variable/scope/Saver machinery the reference scripts use
(``tf.variable_scope``, ``tf.get_variable``, ``tf.train.Saver``; reference
``stylize_image.py:63-73``, ``train.py:158-165,224-225``).

Variables live as numpy arrays in a process-wide store keyed by their TF names
(``img_t_net/initconv_0/W`` ...).  ``Saver.restore`` / ``Saver.save`` read and write the
reference's V2-bundle checkpoints byte-compatibly (``tf_bundle``).
"""
from __future__ import annotations

import contextlib
from collections import OrderedDict

import numpy as np

from . import tf_bundle

_store: "OrderedDict[str, np.ndarray]" = OrderedDict()
_scope: list = []


def reset_default_graph():
    """tf.reset_default_graph() analogue (train.py:155): forget all variables."""
    _store.clear()
    del _scope[:]
    from . import graph
    graph.clear()


@contextlib.contextmanager
def variable_scope(name: str):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


def current_scope() -> str:
    return "/".join(_scope)


def scoped(name: str) -> str:
    s = current_scope()
    return s + "/" + name if s else name


def get_variable(name: str, shape=None, initializer=None) -> np.ndarray:
    """tf.get_variable analogue: returns the stored array, creating it with
    ``initializer(shape)`` on first use."""
    full = scoped(name)
    if full not in _store:
        if initializer is None:
            raise KeyError("variable %s does not exist and no initializer was given" % full)
        v = np.asarray(initializer(tuple(shape)) if callable(initializer) else initializer, np.float32)
        if shape is not None and tuple(v.shape) != tuple(shape):
            raise ValueError("initializer for %s produced shape %s, expected %s" % (full, v.shape, shape))
        _store[full] = v
    elif shape is not None and tuple(_store[full].shape) != tuple(shape):
        raise ValueError("variable %s already exists with shape %s, requested %s"
                         % (full, _store[full].shape, tuple(shape)))
    return _store[full]


def set_variable(full_name: str, value):
    _store[full_name] = np.asarray(value)


def all_variables(scope: str | None = None) -> "OrderedDict[str, np.ndarray]":
    if scope is None:
        return OrderedDict(_store)
    pre = scope.rstrip("/") + "/"
    return OrderedDict((k, v) for k, v in _store.items() if k.startswith(pre))


class Saver:
    """tf.train.Saver analogue over V2-bundle files.

    ``Saver()`` covers every variable currently in the store; ``Saver(var_list)`` only the
    named ones (train.py:224-225)."""

    def __init__(self, var_list=None):
        self._names = None if var_list is None else [v if isinstance(v, str) else v[0] for v in var_list]

    def _target_names(self):
        return list(_store) if self._names is None else list(self._names)

    def restore(self, sess, save_path: str):
        """Load every covered variable from the checkpoint.  Like TF, a variable that is
        missing or has a different shape in the checkpoint is an error (this is how a wrong
        --upsample_method surfaces in the reference, stylize_image.py:37-41)."""
        ck = tf_bundle.read_checkpoint(save_path)
        names = self._target_names()
        if not names:                       # nothing declared yet: adopt the checkpoint's variables
            names = list(ck)
        for n in names:
            if n not in ck:
                raise KeyError("Key %s not found in checkpoint %s" % (n, save_path))
            if n in _store and tuple(_store[n].shape) != tuple(ck[n].shape):
                raise ValueError("Assign requires shapes of both tensors to match. lhs shape= %s rhs shape= %s "
                                 "(variable %s; wrong --upsample_method?)"
                                 % (list(_store[n].shape), list(ck[n].shape), n))
            _store[n] = ck[n]

    def save(self, sess, save_path: str, global_step=None) -> str:
        path = save_path if global_step is None else "%s-%d" % (save_path, int(global_step))
        tf_bundle.write_checkpoint(path, OrderedDict((n, _store[n]) for n in self._target_names()))
        return path
