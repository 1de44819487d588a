"""Input pipeline for train.py (reference datapipe.py:14-78), without TensorFlow.

Sources, chosen from ``train_dir``:
* TFRecord shards ``train-*`` written by the reference's tfrecords_writer.py (framing +
  ``tf.Example`` with ``image/encoded`` JPEG bytes) - decoded with OpenCV;
* a directory of image files (jpg/jpeg/png);
* the literal ``synthetic`` or ``synthetic:<count>`` - seeded uniform-noise images
  (no dataset is reachable from the build machine).
Images are resized to ``resize_shape`` with a restatement of TF-1.0's legacy bicubic kernel
(``tf.image.resize_images(method=2)``, datapipe.py:25) and shuffled through a buffer of
``min_after_dequeue`` elements like ``tf.train.shuffle_batch`` (datapipe.py:71-77).
The reference's TF queue threads are replaced by a plain Python iterator; exhausting
``num_epochs`` raises ``OutOfRangeError`` like the queue does (train.py:281).
"""
from __future__ import annotations

import glob
import os
import struct

import numpy as np


class OutOfRangeError(Exception):
    """Stand-in for tf.errors.OutOfRangeError (train.py:281)."""


# ------------------------------------------------------------------ TF-1.0 bicubic
_TABLE = 1024
_A = -0.75


def _coeff_table():
    # TF evaluates the Keys polynomial in double at the float abscissa and stores floats
    x = (np.arange(_TABLE + 1, dtype=np.float32) / _TABLE).astype(np.float64)
    t = np.empty((_TABLE + 1) * 2, np.float32)
    t[0::2] = ((_A + 2) * x - (_A + 3)) * x * x + 1
    x1 = x + 1
    t[1::2] = ((_A * x1 - 5 * _A) * x1 + 8 * _A) * x1 - 4 * _A
    return t


_TAB = _coeff_table()


def _weights_indices(in_size, out_size):
    scale = np.float32(in_size) / np.float32(out_size)
    out = np.arange(out_size)
    pos = scale * out.astype(np.float32)
    in_loc = np.floor(pos).astype(np.int64)
    delta = pos - in_loc
    off = np.rint(delta * _TABLE).astype(np.int64)
    w = np.stack([_TAB[off * 2 + 1], _TAB[off * 2], _TAB[(_TABLE - off) * 2], _TAB[(_TABLE - off) * 2 + 1]], 1)
    idx = np.stack([in_loc - 1, in_loc, in_loc + 1, in_loc + 2], 1).clip(0, in_size - 1)
    return w.astype(np.float32), idx


def resize_bicubic_tf1(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """TF-1.0 ResizeBicubic (align_corners=False, no half-pixel centres, A=-0.75, 1024-entry
    coefficient table), HWC uint8/float -> HWC float32.  Unpinned restatement."""
    img = np.asarray(img, np.float32)
    wy, iy = _weights_indices(img.shape[0], out_h)
    wx, ix = _weights_indices(img.shape[1], out_w)
    # TF's order: per output pixel four horizontal interpolations (one per patch row), then one vertical;
    # products summed left to right in float32
    acc = None
    for i in range(4):
        rows = img[iy[:, i]]                                         # [out_h, W, C]
        r = None
        for k in range(4):
            pr = rows[:, ix[:, k]] * wx[None, :, k, None]
            r = pr if r is None else r + pr
        pr = r * wy[:, i, None, None]
        acc = pr if acc is None else acc + pr
    return acc.astype(np.float32)


# ------------------------------------------------------------------ TFRecord / tf.Example
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
            v = b[p:p + 4]; p += 4
        elif wt == 1:
            v = b[p:p + 8]; p += 8
        else:
            raise ValueError("bad wire type")
        yield f, v


def iter_tfrecord(path):
    """Yield raw records of a TFRecord file (length:u64, crc:u32, data, crc:u32)."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if len(head) < 12:
                return
            n = struct.unpack("<Q", head[:8])[0]
            data = f.read(n)
            f.read(4)
            if len(data) < n:
                return
            yield data


def example_bytes_feature(record: bytes, key: str):
    """Extract a bytes feature of a serialized tf.Example (features{feature{key,value}})."""
    for f, v in _fields(record):
        if f != 1:
            continue
        for f2, entry in _fields(v):              # map entries
            if f2 != 1:
                continue
            k, val = None, None
            for f3, v3 in _fields(entry):
                if f3 == 1:
                    k = bytes(v3).decode()
                elif f3 == 2:
                    val = v3
            if k == key and val is not None:
                for f4, v4 in _fields(val):        # Feature: 1 = bytes_list
                    if f4 == 1:
                        for f5, v5 in _fields(v4):
                            if f5 == 1:
                                return bytes(v5)
    return None


# ------------------------------------------------------------------ batcher
def _image_stream(train_dir, num_epochs, order_rng, img_rng, shard=(0, 1)):
    """Decoded RGB uint8 images of this rank's shard.  The stream position i counts EVERY record of the
    multi-epoch stream (decodable or not) and only records with i % world == rank are decoded - the other ranks'
    records cost one framing read (TFRecord) or nothing (image files), not a JPEG decode.  ``order_rng`` drives
    only the per-epoch shard / file order and must be seeded identically on every rank (the shards of the ranks
    are disjoint only if all ranks walk the same order); ``img_rng`` is rank-private (synthetic images)."""
    import cv2
    rank, world = shard
    i = 0
    if train_dir.startswith("synthetic"):
        count = int(train_dir.split(":")[1]) if ":" in train_dir else 1 << 30
        for _ in range(num_epochs if num_epochs else 1 << 30):
            for _ in range(count):
                mine = i % world == rank
                i += 1
                if mine:
                    yield img_rng.randint(0, 256, (256, 256, 3)).astype(np.uint8)
        return
    shards = sorted(glob.glob(os.path.join(train_dir, "train-*")))
    files = [p for p in sorted(glob.glob(os.path.join(train_dir, "*")))
             if p.lower().endswith((".jpg", ".jpeg", ".png"))]
    if not shards and not files:
        raise IOError("no TFRecord shards (train-*) or image files found in %r" % train_dir)
    for _ in range(num_epochs if num_epochs else 1 << 30):
        if shards:
            order = list(shards); order_rng.shuffle(order)         # string_input_producer(shuffle=True)
            for s in order:
                for rec in iter_tfrecord(s):
                    mine = i % world == rank
                    i += 1
                    if not mine:
                        continue
                    enc = example_bytes_feature(rec, "image/encoded")
                    if enc is None:
                        continue
                    img = cv2.imdecode(np.frombuffer(enc, np.uint8), cv2.IMREAD_COLOR)
                    if img is not None:
                        yield cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        else:
            order = list(files); order_rng.shuffle(order)
            for p in order:
                mine = i % world == rank
                i += 1
                if not mine:
                    continue
                img = cv2.imread(p)
                if img is not None:
                    yield cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def next_batch_collective(batches, group=None):
    """``next(batches)`` with a COLLECTIVE end of data: every rank of ``group`` (a torch.distributed group that
    accepts CPU tensors, e.g. gloo) votes has-batch with an all-reduce(MIN); if any rank's pipeline is exhausted,
    ALL ranks raise OutOfRangeError at this same step, so no rank is left waiting in the next gradient all-reduce
    (ranks can hold different batch counts: shards differ by one image and undecodable files drop per rank).
    ``group=None``: plain single-process behaviour."""
    have = True
    batch = None
    try:
        batch = next(batches)
    except (OutOfRangeError, StopIteration):
        have = False
    if group is not None:
        import torch
        import torch.distributed as dist
        vote = torch.tensor([1 if have else 0], dtype=torch.int32)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=group)
        have = bool(int(vote.item()))
    if not have:
        raise OutOfRangeError("input pipeline exhausted" + ("" if group is None else " on at least one rank"))
    return batch


class Prefetcher:
    """The reference's input pipeline runs on TF queue-runner threads beside the training loop
    (train.py:241-242, shuffle_batch capacity = min_after_dequeue + 3 batches, datapipe.py:72-77).  This is the same
    producer/consumer split without TensorFlow: one daemon thread drives the batch iterator (TFRecord parsing, JPEG
    decode - OpenCV releases the GIL -, resize, shuffle) into a bounded queue of ``depth`` batches; the consumer
    sees the batches in the producer's order and the producer's terminal exception (OutOfRangeError)."""

    _END = object()

    def __init__(self, iterator, depth=3):
        import queue
        import threading
        self._q = queue.Queue(maxsize=max(1, int(depth)))
        self._it = iterator
        self._exc = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, name="faststyle-datapipe", daemon=True)
        self._t.start()

    def _put(self, item):
        import queue
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _run(self):
        try:
            for item in self._it:
                if not self._put(item):
                    return
        except BaseException as e:          # noqa: BLE001 - handed to the consumer
            self._exc = e
        self._put(self._END)

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._END:
            self._q.put(self._END)          # stay terminated for later calls
            if self._exc is not None:
                raise self._exc
            raise StopIteration
        return item

    def close(self):
        self._stop.set()


def prefetch(iterator, depth=3):
    return Prefetcher(iterator, depth)


class GpuPreprocessor:
    """datapipe.preprocessing (reference datapipe.py:14-26) on the device: the host only decodes; each uint8 image
    goes up through a pinned staging buffer and `fs_resize_bicubic_tf1_u8` writes it, resized and cast to float32,
    straight into its slot of the [B, h, w, 3] batch tensor the train step reads (no float32 batch on the host,
    a quarter of the H2D bytes for same-size images, far fewer for larger ones)."""

    def __init__(self, batch_size, resize_shape, device="cuda:0"):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.h, self.w = int(resize_shape[0]), int(resize_shape[1])
        self.batch = torch.empty((int(batch_size), self.h, self.w, 3), dtype=torch.float32, device=self.device)
        self._stage = None

    def __call__(self, images):
        """images: sequence of HWC uint8 RGB arrays (any sizes), len == batch_size -> the device batch tensor."""
        from .ops import resize_bicubic_tf1
        torch = self.torch
        if len(images) != self.batch.shape[0]:
            raise ValueError("expected %d images, got %d" % (self.batch.shape[0], len(images)))
        with torch.cuda.device(self.device):
            for i, img in enumerate(images):
                img = np.ascontiguousarray(img, dtype=np.uint8)
                n = img.size
                if self._stage is None or self._stage.numel() < n:
                    self._stage = torch.empty(max(n, 1 << 20), dtype=torch.uint8).pin_memory()
                    self._dev_stage = torch.empty(self._stage.numel(), dtype=torch.uint8, device=self.device)
                    self._free = torch.cuda.Event()
                else:
                    self._free.synchronize()            # the previous image has left the staging buffer
                self._stage[:n].copy_(torch.from_numpy(img.reshape(-1)))
                src = self._dev_stage[:n]
                src.copy_(self._stage[:n], non_blocking=True)
                self._free.record()
                resize_bicubic_tf1(src.view(img.shape), self.h, self.w, out=self.batch[i])
        return self.batch


def batcher(train_dir, batch_size, resize_shape=None, num_epochs=None, min_after_dequeue=4000, seed=0,
            shard=(0, 1), raw=False):
    """Iterator of float32 NHWC batches.  ``shard=(rank, world)`` keeps every world-th image for
    data-parallel training (sharded BEFORE decoding).  Raises OutOfRangeError when the epochs are exhausted;
    ranks may reach that point one batch apart - pair with ``next_batch_collective``.
    ``raw=True`` yields lists of decoded uint8 images instead (un-resized), for GpuPreprocessor."""
    # one RNG, seeded identically on every rank, ONLY for the epoch order; a rank-private one for everything whose
    # consumption depends on the rank's own stream position (shuffle-buffer sampling, synthetic images)
    order_rng = np.random.RandomState(seed)
    rng = np.random.RandomState((seed * 9973 + 7919 * (shard[0] + 1)) % (1 << 31))
    stream = _image_stream(train_dir, num_epochs, order_rng, rng, shard)
    buf = []

    def prep(img):
        if raw:
            return np.ascontiguousarray(img, dtype=np.uint8)
        if resize_shape is None:
            return np.asarray(img, np.float32)
        if tuple(img.shape[:2]) == tuple(resize_shape):
            return np.asarray(img, np.float32)
        return resize_bicubic_tf1(img, resize_shape[0], resize_shape[1])

    def gen():
        exhausted = False
        i = 0
        while True:
            while not exhausted and len(buf) < min_after_dequeue + batch_size:
                try:
                    img = next(stream)
                except StopIteration:
                    exhausted = True
                    break
                buf.append(prep(img))
            if len(buf) < batch_size:
                raise OutOfRangeError("input pipeline exhausted")
            out = []
            for _ in range(batch_size):
                j = rng.randint(len(buf))
                buf[j], buf[-1] = buf[-1], buf[j]
                out.append(buf.pop())
            yield out if raw else np.stack(out, 0)
    return gen()
