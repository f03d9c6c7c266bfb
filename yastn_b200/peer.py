"""Peer arenas: every rank's operand buffer mapped into the address space of the other ranks of the box.

One process per GPU.  ``PeerArena`` allocates a device buffer through the C ABI (``yb_peer_alloc``), ships its 64-byte
CUDA IPC handle to the other ranks with ``torch.distributed.all_gather_object`` and maps theirs (``yb_peer_open``).
Tensors carved out of the arena sit at the SAME offset on every rank, so "block [lo, hi) of tensor X on rank r" is the
address ``peer_ptr[r] + offset(X) + lo * itemsize`` — known on every rank from metadata alone.  The block exchange between
two sharded contractions is then one launch of the block-copy kernel whose records point into the peers' buffers
(``sharding.PeerExchange``), or no extra launch at all when the producing grouped GEMM scatters its result blocks straight
to their next owners (``backend_b200.dot_unmerge(..., out=, dst_shift=)``).  ``publish()`` is the only collective: a
one-element all-reduce on the caller's stream, after which every block written by a peer is visible locally.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

class _Span:
    """A window of device memory (bytes) exposed through the CUDA array interface (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self.owner = owner      # keeps the arena alive as long as a tensor views it


class PeerArena:
    def __init__(self, nbytes, device=None, group=None):
        self._lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nbytes = int(nbytes)
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(self._lib.yb_peer_alloc(self.nbytes, self.device.index, ctypes.byref(ptr), handle))
        self.ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.device.index, bytes(handle)), group=group)
        self.peer_ptrs = []
        for r, (dev_r, h) in enumerate(handles):
            if r == self.rank:
                self.peer_ptrs.append(self.ptr)
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            _lib.check(self._lib.yb_peer_open(buf, self.device.index, ctypes.byref(p)))
            self.peer_ptrs.append(p.value)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._top = 0
        self._closed = False

    # ---- carving tensors out of the arena (same sequence of calls on every rank => same offsets everywhere)
    def empty(self, numel, dtype):
        isz = torch.empty(0, dtype=dtype).element_size()
        off = (self._top + 255) // 256 * 256
        if off + numel * isz > self.nbytes:
            raise MemoryError(f"PeerArena: {numel * isz} bytes requested, {self.nbytes - off} left")
        self._top = off + numel * isz
        return self.view(off, numel, dtype)

    def view(self, offset_bytes, numel, dtype):
        if numel == 0:
            return torch.empty(0, dtype=dtype, device=self.device)
        isz = torch.empty(0, dtype=dtype).element_size()
        return torch.as_tensor(_Span(self.ptr + offset_bytes, numel * isz, self), device=self.device).view(dtype)

    def reset(self):
        self._top = 0

    def shift(self, peer, itemsize):
        """Element offset that turns an address inside the local arena into the same address inside ``peer``'s arena."""
        d = self.peer_ptrs[peer] - self.ptr
        if d % itemsize:
            raise ValueError("peer arena mappings are not aligned to the element size")
        return d // itemsize

    def contains(self, t):
        return self.ptr <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= self.ptr + self.nbytes

    def publish(self):
        """Stream-ordered: returns at once; work queued afterwards on the current stream sees every block that any rank
        pushed before ITS publish()."""
        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)       # nobody may still be writing into a buffer that is about to disappear
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                self._lib.yb_peer_close(ctypes.c_void_p(p))
        dist.barrier(group=self.group)
        self._lib.yb_peer_free(ctypes.c_void_p(self.ptr))
