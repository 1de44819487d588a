"""G3: the V2-bundle reader/writer is byte-exact on the reference's shipped checkpoints
(reference models/*.ckpt; written by tf.train.Saver at train.py:286)."""
import os

import numpy as np
import pytest

from faststyle_b200 import tf_bundle as tb
from faststyle_b200.layout import TRANSFORM_VARS, flatten_transform, unflatten_transform
from oracle import ckpt as ockpt


@pytest.mark.parametrize("model", ["starry", "candy"])
def test_roundtrip_is_byte_identical(golden_dir, model):
    prefix = os.path.join(golden_dir, model + "_final.ckpt")
    ck = tb.read_checkpoint(prefix)               # verifies block + tensor CRC32Cs
    assert [(k, v.shape) for k, v in ck.items()] == [(n, tuple(s)) for n, s in TRANSFORM_VARS]
    idx, data = tb.serialize_checkpoint(ck)
    assert idx == open(prefix + ".index", "rb").read()
    assert data == open(prefix + ".data-00000-of-00001", "rb").read()


def test_independent_parsers_agree(golden_dir):
    prefix = os.path.join(golden_dir, "starry_final.ckpt")
    a, b = tb.read_checkpoint(prefix), ockpt.load(prefix)
    assert list(a) == list(b)
    for k in a:
        assert np.array_equal(a[k], b[k])


def test_write_read_and_errors(tmp_path, golden_dir):
    ck = tb.read_checkpoint(os.path.join(golden_dir, "candy_final.ckpt"))
    ck["global_step"] = np.array(7, dtype=np.int64)            # scalar, other dtype
    ck["zz/empty_dim"] = np.zeros((0, 3), np.float32)
    tb.write_checkpoint(str(tmp_path / "m.ckpt"), ck)
    back = tb.read_checkpoint(str(tmp_path / "m.ckpt"))
    assert back["global_step"].shape == () and int(back["global_step"]) == 7 and back["zz/empty_dim"].shape == (0, 3)
    for k in ck:
        assert np.array_equal(ck[k], back[k])
    assert "model_checkpoint_path" in open(tmp_path / "checkpoint").read()
    # corruption is detected
    p = str(tmp_path / "m.ckpt.data-00000-of-00001")
    raw = bytearray(open(p, "rb").read()); raw[100] ^= 0xFF
    open(p, "wb").write(raw)
    with pytest.raises(ValueError, match="CRC"):
        tb.read_checkpoint(str(tmp_path / "m.ckpt"))
    with pytest.raises(FileNotFoundError):
        tb.read_checkpoint(str(tmp_path / "missing.ckpt"))


def test_many_entries_multi_block(tmp_path):
    """> 256 KiB of index entries forces several data blocks + separator keys."""
    rng = np.random.RandomState(0)
    tensors = {"scope_%05d/some/rather/long/variable/name/to/fill/blocks_%d" % (i, i):
               rng.standard_normal((2, 3)).astype(np.float32) for i in range(4000)}
    tb.write_checkpoint(str(tmp_path / "big.ckpt"), tensors)
    back = tb.read_checkpoint(str(tmp_path / "big.ckpt"))
    assert len(back) == 4000 and list(back) == sorted(tensors)
    for k in tensors:
        assert np.array_equal(tensors[k], back[k])


def test_flat_layout_roundtrip(golden_dir):
    ck = tb.read_checkpoint(os.path.join(golden_dir, "starry_final.ckpt"))
    flat = flatten_transform(ck)
    assert flat.tobytes() == open(os.path.join(golden_dir, "starry_final.ckpt.data-00000-of-00001"), "rb").read()
    back = unflatten_transform(flat)
    assert all(np.array_equal(ck[k], back[k]) for k in ck)
    bad = dict(ck); bad["img_t_net/upsample_0/W"] = np.zeros((3, 3, 32, 64), np.float32)   # a 'deconv' model
    with pytest.raises(ValueError, match="upsample_method"):
        flatten_transform(bad)
