"""TensorFlow "V2 bundle" checkpoint reader / writer, without TensorFlow.

The reference saves and restores its transform-net weights with
``tf.train.Saver`` (reference ``train.py:224-225,259,286``,
``stylize_image.py:68-73``).  That produces two files,

* ``<prefix>.index``              – a LevelDB-format SSTable whose values are
                                    ``BundleEntryProto`` messages, and
* ``<prefix>.data-00000-of-00001`` – the raw little-endian tensor bytes,
                                    concatenated in byte-sorted key order.

This module parses and emits exactly that layout so that files written here
load in stock TensorFlow and the shipped ``models/*.ckpt`` load here.  The
writer is byte-exact: re-serialising the shipped checkpoints reproduces the
original ``.index`` and ``.data`` files (tests/test_tf_bundle.py).

Layout facts were decoded from the shipped files.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

_MAGIC = bytes.fromhex("57fb808b247547db")
_MASK_DELTA = 0xA282EAD8
_RESTART_INTERVAL = 16
_BLOCK_SIZE = 262144          # TF table_options.h default

# TF DataType enum values we support.
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_DT_TO_NP = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"),
             DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_NP_TO_DT = {v: k for k, v in _DT_TO_NP.items()}


# --------------------------------------------------------------------------
# CRC32C (Castagnoli), table driven, numpy-accelerated over 8-byte strides
# --------------------------------------------------------------------------
def _make_tables():
    poly = 0x82F63B78
    t0 = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t0[i] = c
    tabs = [t0]
    for _ in range(7):
        prev = tabs[-1]
        tabs.append((prev >> 8) ^ t0[prev & 0xFF])
    return [t.tolist() for t in tabs]


_T = _make_tables()


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC32C of ``data`` (slicing-by-8)."""
    c = crc ^ 0xFFFFFFFF
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    n = len(data)
    i = 0
    mv = memoryview(data)
    end8 = n - (n % 8)
    if end8:
        words = struct.unpack_from("<%dQ" % (end8 // 8), mv, 0)
        for w in words:
            w ^= c
            c = (t7[w & 0xFF] ^ t6[(w >> 8) & 0xFF] ^ t5[(w >> 16) & 0xFF] ^
                 t4[(w >> 24) & 0xFF] ^ t3[(w >> 32) & 0xFF] ^
                 t2[(w >> 40) & 0xFF] ^ t1[(w >> 48) & 0xFF] ^ t0[(w >> 56) & 0xFF])
        i = end8
    while i < n:
        c = t0[(c ^ mv[i]) & 0xFF] ^ (c >> 8)
        i += 1
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# --------------------------------------------------------------------------
# varint / protobuf helpers (only what BundleHeader/BundleEntry need)
# --------------------------------------------------------------------------
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _parse_proto(buf: bytes):
    """Yield (field_number, wire_type, value) for a flat protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _encode_entry(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    """BundleEntryProto{1:dtype 2:shape 4:offset 5:size 6:crc32c}; proto3
    omits zero-valued scalars (the first entry has no offset field)."""
    out = bytearray()
    if dtype:
        out += b"\x08" + _put_varint(dtype)
    shp = bytearray()
    for d in shape:
        dim = (b"\x08" + _put_varint(int(d))) if d else b""
        shp += b"\x12" + _put_varint(len(dim)) + dim
    out += b"\x12" + _put_varint(len(shp)) + bytes(shp)
    if offset:
        out += b"\x20" + _put_varint(offset)
    if size:
        out += b"\x28" + _put_varint(size)
    if crc:
        out += b"\x35" + struct.pack("<I", crc)
    return bytes(out)


def _decode_entry(buf: bytes):
    dtype, shape, shard, offset, size, crc = 0, [], 0, 0, 0, 0
    for f, wt, v in _parse_proto(buf):
        if f == 1:
            dtype = v
        elif f == 2:
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:            # dim
                    sz = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            sz = v3
                    shape.append(sz)
        elif f == 3:
            shard = v
        elif f == 4:
            offset = v
        elif f == 5:
            size = v
        elif f == 6:
            crc = v
    return dict(dtype=dtype, shape=tuple(shape), shard_id=shard,
                offset=offset, size=size, crc32c=crc)


# --------------------------------------------------------------------------
# SSTable blocks
# --------------------------------------------------------------------------
def _read_block(buf: bytes, offset: int, size: int, verify=True) -> bytes:
    block = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if trailer[0] != 0:
        raise ValueError("compressed SSTable blocks are not supported")
    if verify:
        want = struct.unpack("<I", trailer[1:5])[0]
        got = masked_crc32c(block + trailer[:1])
        if want != got:
            raise ValueError("SSTable block CRC mismatch at offset %d" % offset)
    return block


def _iter_block(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        val = block[pos:pos + vlen]
        pos += vlen
        yield key, val


class _BlockBuilder:
    def __init__(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""

    def add(self, key: bytes, val: bytes):
        shared = 0
        if self.counter < _RESTART_INTERVAL:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val))
        self.buf += key[shared:] + val
        self.last_key = key
        self.counter += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self) -> bytes:
        out = bytes(self.buf)
        out += b"".join(struct.pack("<I", r) for r in self.restarts)
        out += struct.pack("<I", len(self.restarts))
        return out


def _shortest_separator(a: bytes, b: bytes) -> bytes:
    """LevelDB BytewiseComparator::FindShortestSeparator."""
    m = min(len(a), len(b))
    i = 0
    while i < m and a[i] == b[i]:
        i += 1
    if i < m and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def _short_successor(a: bytes) -> bytes:
    """LevelDB BytewiseComparator::FindShortSuccessor."""
    for i, c in enumerate(a):
        if c != 0xFF:
            return a[:i] + bytes([c + 1])
    return a


# --------------------------------------------------------------------------
# public API
# --------------------------------------------------------------------------
def _data_path(prefix: str, shard: int = 0, n: int = 1) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, n)


def read_index(prefix: str, verify=True) -> "OrderedDict[str, dict]":
    """Parse ``<prefix>.index`` → ordered {name: entry dict}."""
    path = prefix + ".index"
    if not os.path.exists(path):
        raise FileNotFoundError(
            "checkpoint index %r not found (expected a TF V2 bundle: "
            "<prefix>.index + <prefix>.data-00000-of-00001)" % path)
    buf = open(path, "rb").read()
    if len(buf) < 48 or buf[-8:] != _MAGIC:
        raise ValueError("%s is not an SSTable (bad magic)" % path)
    footer = buf[-48:]
    pos = 0
    _mo, pos = _get_varint(footer, pos)
    _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    index_block = _read_block(buf, io, isz, verify)
    entries = OrderedDict()
    header_seen = False
    for _, handle in _iter_block(index_block):
        off, p = _get_varint(handle, 0)
        sz, p = _get_varint(handle, p)
        for key, val in _iter_block(_read_block(buf, off, sz, verify)):
            if key == b"":
                header_seen = True
                hdr = {f: v for f, _, v in _parse_proto(val)}
                if hdr.get(1, 1) != 1:
                    raise ValueError("multi-shard bundles are not supported")
                continue
            entries[key.decode("utf-8")] = _decode_entry(val)
    if not header_seen:
        raise ValueError("%s has no bundle header entry" % path)
    return entries


def read_checkpoint(prefix: str, verify=True) -> "OrderedDict[str, np.ndarray]":
    """Load every tensor of a V2 bundle as numpy arrays (name → array)."""
    entries = read_index(prefix, verify)
    with open(_data_path(prefix), "rb") as f:
        data = f.read()
    out = OrderedDict()
    for name, e in entries.items():
        if e["dtype"] not in _DT_TO_NP:
            raise ValueError("tensor %s has unsupported dtype enum %d" % (name, e["dtype"]))
        raw = data[e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError("tensor %s is truncated in the data shard" % name)
        if verify and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError("tensor %s failed its CRC32C check" % name)
        out[name] = np.frombuffer(raw, dtype=_DT_TO_NP[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def serialize_checkpoint(tensors) -> "tuple[bytes, bytes]":
    """Return (index_bytes, data_bytes) for {name: ndarray}."""
    def _c(v):                       # C-contiguous without promoting 0-d scalars to 1-d
        a = np.asarray(v)
        return a if a.flags.c_contiguous else np.ascontiguousarray(a)
    items = sorted(((k.encode("utf-8"), _c(v)) for k, v in tensors.items()), key=lambda kv: kv[0])
    data = bytearray()
    kvs = [(b"", b"\x08\x01\x1a\x02\x08\x01")]   # BundleHeaderProto{num_shards=1, version{producer=1}}
    for key, arr in items:
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        dt = np.dtype(dt.str.replace("=", "<").replace("|", "<")) if dt.str[0] in "=|" else dt
        if dt not in _NP_TO_DT:
            raise ValueError("unsupported dtype %s for tensor %s" % (arr.dtype, key))
        raw = arr.astype(dt, copy=False).tobytes()
        kvs.append((key, _encode_entry(_NP_TO_DT[dt], arr.shape, len(data), len(raw),
                                       masked_crc32c(raw))))
        data += raw

    out = bytearray()
    index = _BlockBuilder()

    def emit(block_bytes: bytes):
        off = len(out)
        out.extend(block_bytes)
        out.extend(b"\x00" + struct.pack("<I", masked_crc32c(block_bytes + b"\x00")))
        return off, len(block_bytes)

    blk = _BlockBuilder()
    pending = None                      # (last_key, handle) of a flushed block
    for key, val in kvs:
        if pending is not None:
            sep = _shortest_separator(pending[0], key)
            index.add(sep, pending[1])
            pending = None
        blk.add(key, val)
        if blk.size_estimate() >= _BLOCK_SIZE:
            off, sz = emit(blk.finish())
            pending = (blk.last_key, _put_varint(off) + _put_varint(sz))
            blk = _BlockBuilder()
    if not blk.empty():
        off, sz = emit(blk.finish())
        pending = (blk.last_key, _put_varint(off) + _put_varint(sz))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    moff, msz = emit(_BlockBuilder().finish())          # empty metaindex block
    ioff, isz = emit(index.finish())
    footer = _put_varint(moff) + _put_varint(msz) + _put_varint(ioff) + _put_varint(isz)
    footer += b"\x00" * (40 - len(footer)) + _MAGIC
    out.extend(footer)
    return bytes(out), bytes(data)


def write_checkpoint(prefix: str, tensors) -> None:
    """Write {name: ndarray} as ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``
    and update the sibling ``checkpoint`` state file like ``tf.train.Saver.save``."""
    index_bytes, data_bytes = serialize_checkpoint(tensors)
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(_data_path(prefix), "wb") as f:
        f.write(data_bytes)
    with open(prefix + ".index", "wb") as f:
        f.write(index_bytes)
    base = os.path.basename(prefix)
    with open(os.path.join(d or ".", "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
