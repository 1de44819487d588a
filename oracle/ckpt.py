"""Independent minimal reader for the reference's TF V2-bundle checkpoints
(TEST ORACLE — see oracle/__init__.py).  Deliberately shares no code with
``faststyle_b200.tf_bundle`` so the two parsers check each other.

Relies on two layout facts decoded from the shipped ``models/*.ckpt`` files
(SURVEY.md App. B): index entries are prefix-compressed (shared, non_shared,
value_len varints) inside one uncompressed data block, and each value is a
BundleEntryProto {1:dtype 2:shape 4:offset 5:size 6:crc}.
"""
import struct
import numpy as np


def _varint(b, p):
    v = s = 0
    while True:
        c = b[p]; p += 1
        v |= (c & 127) << s; s += 7
        if c < 128:
            return v, p


def _fields(b):
    p = 0
    while p < len(b):
        key, p = _varint(b, p)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _varint(b, p)
        elif wt == 2:
            n, p = _varint(b, p); v = b[p:p + n]; p += n
        elif wt == 5:
            v = struct.unpack_from("<I", b, p)[0]; p += 4
        else:
            raise ValueError("wire type %d" % wt)
        yield f, v


def load(prefix):
    """Return {name: float32 ndarray} for a single-shard, single-data-block bundle."""
    idx = open(prefix + ".index", "rb").read()
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert idx[-8:] == bytes.fromhex("57fb808b247547db"), "not an SSTable"
    foot = idx[-48:]
    p = 0
    _, p = _varint(foot, p); _, p = _varint(foot, p)
    ioff, p = _varint(foot, p); isz, p = _varint(foot, p)
    iblk = idx[ioff:ioff + isz]
    # single index entry: shared, non_shared, vlen, key, handle(offset,size)
    q = 0
    _, q = _varint(iblk, q); ns, q = _varint(iblk, q); _, q = _varint(iblk, q); q += ns
    doff, q = _varint(iblk, q); dsz, q = _varint(iblk, q)
    blk = idx[doff:doff + dsz]
    nrest = struct.unpack_from("<I", blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * nrest
    out, key, p = {}, b"", 0
    while p < end:
        sh, p = _varint(blk, p); ns, p = _varint(blk, p); vl, p = _varint(blk, p)
        key = key[:sh] + blk[p:p + ns]; p += ns
        val = blk[p:p + vl]; p += vl
        if not key:
            continue
        shape, off, size = [], 0, 0
        for f, v in _fields(val):
            if f == 1:
                assert v == 1, "only DT_FLOAT"
            elif f == 2:
                for f2, v2 in _fields(v):
                    d = 0
                    for f3, v3 in _fields(v2):
                        d = v3
                    shape.append(d)
            elif f == 4:
                off = v
            elif f == 5:
                size = v
        out[key.decode()] = np.frombuffer(data[off:off + size], "<f4").reshape(shape).copy()
    return out
