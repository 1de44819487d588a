"""Peer-memory exchange buffer for the data-parallel optimiser step.

One process per GPU on one node: every rank allocates the same flat float32 buffer (``fs_peer_buffer_create``: a plain
device allocation + its CUDA IPC handle), the 64-byte handles travel through ``torch.distributed`` (plumbing only), and
every rank maps the buffers of all peers into its own device's address space (``fs_peer_buffer_open``; the open enables
NVLink peer access).  ``fs_dp_allreduce_adam`` - ONE kernel: gradient all-reduce(SUM) by peer loads + TF-Adam - then
reads every rank's gradients directly.

Layout of every rank's buffer (floats): two parity halves ``[n gradients | pad | n_extra scalars | pad]`` (step t uses
half t & 1: a rank can be at most one step ahead of the slowest peer, so a half is never rewritten while a peer still
reads it), then 64 uint32 flag slots owned by the kernel."""
from __future__ import annotations

import ctypes as C

import torch


class _DevArray:
    """Minimal __cuda_array_interface__ carrier: lets torch view a raw device allocation without copying."""

    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}


class PeerExchange:
    def __init__(self, n: int, n_extra: int, device, group):
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("PeerExchange: at most 16 ranks (one NVLink domain)")
        self.n, self.n_extra = int(n), int(n_extra)
        self.extra_off = (self.n + 7) // 8 * 8
        self.stride = (self.extra_off + self.n_extra + 63) // 64 * 64
        self.flag_off_bytes = 2 * self.stride * 4
        total = 2 * self.stride + 64
        self.device = torch.device(device)
        self._opened = []
        self._base = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _lib.call("fs_peer_buffer_create", C.c_size_t(total * 4), C.byref(self._base), C.cast(handle, C.c_void_p))
            self.buf = torch.as_tensor(_DevArray(self._base.value, total), device=self.device)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            ptrs = []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    ptrs.append(self._base.value)
                    continue
                p = C.c_void_p()
                hbuf = (C.c_ubyte * 64).from_buffer_copy(hb)
                _lib.call("fs_peer_buffer_open", C.cast(hbuf, C.c_void_p), C.byref(p))
                self._opened.append(p.value)
                ptrs.append(p.value)
        self.ptr_array = (C.c_void_p * self.world)(*ptrs)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.extra_sum = torch.zeros(max(self.n_extra, 1), dtype=torch.float32, device=self.device)
        dist.barrier(group)                      # every buffer is zeroed and mapped before the first kernel runs

    def close(self):
        """Unmap the peers' buffers and free ours (collective in spirit: call it on every rank after the last step)."""
        if getattr(self, "_lib", None) is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                self._lib.call("fs_peer_buffer_close", C.c_void_p(p))
            self._opened = []
            if self._base.value:
                self.buf = None
                self._lib.call("fs_peer_buffer_free", self._base)
                self._base = C.c_void_p()

    def grads(self, parity: int) -> torch.Tensor:
        o = (parity & 1) * self.stride
        return self.buf[o:o + self.n]

    def extra(self, parity: int) -> torch.Tensor:
        o = (parity & 1) * self.stride + self.extra_off
        return self.buf[o:o + self.n_extra]

    def allreduce_adam(self, opt, parity: int, tag: int):
        """All-reduce(SUM) of every rank's gradients of this step + TF-Adam on ``opt``'s buffers; the summed extra
        scalars land in ``self.extra_sum``.  One kernel launch on the current stream.  ``tag`` = the step number
        (1, 2, ...), or 0 to let the kernel take ``opt.step_counter + 1`` on the device (CUDA-graph replay)."""
        from .engine import ptr, stream_ptr
        with torch.cuda.device(self.device):
            self._lib.call("fs_dp_allreduce_adam", C.cast(self.ptr_array, C.POINTER(C.c_void_p)), self.rank, self.world,
                           C.c_longlong((parity & 1) * self.stride), C.c_longlong(self.n), C.c_longlong(self.extra_off),
                           self.n_extra, C.c_longlong(self.flag_off_bytes), C.c_uint(tag & 0xFFFFFFFF), ptr(opt.p),
                           ptr(opt.m), ptr(opt.v), opt.lr, opt.b1, opt.b2, opt.eps, ptr(opt.step_counter),
                           ptr(self.extra_sum), ptr(self.err), stream_ptr())

    def check(self):
        """Raise if a peer's flag never arrived (host-synchronising; call it off the hot path)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("fs_dp_allreduce_adam: a peer rank did not reach the step within 10 s")
