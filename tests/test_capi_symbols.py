"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol
that include/faststyle_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "faststyle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n


def test_binding_covers_header(built_lib):
    from faststyle_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_functions()
    lib = _lib.load()
    assert lib.fs_version() == 100
    assert lib.fs_transform_param_count() == 424102
    assert lib.fs_vgg_flat_floats() == 7635264


def test_layout_queries_match_checkpoint_order(built_lib):
    import numpy as np
    from faststyle_b200 import _lib
    from faststyle_b200.layout import transform_offsets
    lib = _lib.load()
    offs = transform_offsets()
    kinds = {0: "W", 1: "INscale", 2: "INshift"}

    def name(i, w):
        if i < 3:
            return "initconv_%d/%s" % (i, kinds[w])
        if i >= 13:
            return "upsample_%d/%s" % (i - 13, kinds[w])
        return "resblock_%d/%s%d" % ((i - 3) // 2, kinds[w], (i - 3) % 2 + 1)
    for i in range(16):
        for w in range(3):
            o, c = ctypes.c_longlong(), ctypes.c_longlong()
            assert lib.fs_transform_param_slot(i, w, ctypes.byref(o), ctypes.byref(c)) == 0
            eo, es = offs["img_t_net/" + name(i, w)]
            assert (o.value, c.value) == (eo, int(np.prod(es)))


def test_errors_are_reported_without_gpu(built_lib):
    from faststyle_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.fs_engine_create(1, 20, 20, 1, 0, 0, ctypes.byref(h)) != 0       # H <= 40: reflect pad impossible
    assert b"reflect" in lib.fs_last_error()
    assert lib.fs_engine_create(1, 64, 64, 2, 0, 0, ctypes.byref(h)) != 0       # BWD without TRANSFORM
    assert lib.fs_engine_create(2, 256, 256, 15, 1 << 6, 0b1001001010, ctypes.byref(h)) == 0
    assert lib.fs_engine_workspace_bytes(h) > 100e6
    oh, ow = ctypes.c_int(), ctypes.c_int()
    lib.fs_engine_output_dims(h, ctypes.byref(oh), ctypes.byref(ow))
    assert (oh.value, ow.value) == (256, 256)
    lib.fs_engine_destroy(h)
    assert lib.fs_engine_create(1, 474, 712, 1, 0, 0, ctypes.byref(h)) == 0
    lib.fs_engine_output_dims(h, ctypes.byref(oh), ctypes.byref(ow))
    assert (oh.value, ow.value) == (476, 712)
    lib.fs_engine_destroy(h)


def test_product_does_not_import_oracle():
    """The product package and CLIs must never route through oracle/."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "faststyle_b200")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(d, f)).read(), re.M):
                bad.append(f)
    for f in ("stylize_image.py", "train.py", "slow_style.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p) and re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), re.M):
            bad.append(f)
    assert not bad, bad
