"""Symmetric heap: one segment per rank, every segment mapped in every process at ``base + r * stride``.

GPU backend  : CUDA VMM (cuMemCreate + POSIX-fd export/import + optional NVLS multicast), csrc/runtime/symm_heap.cu
CPU backend  : POSIX shared memory, csrc/host/td_host.cpp  (protocol emulation under gloo, no GPU needed)

Replaces the reference's NVSHMEM symmetric heap (``nvshmem.core.tensor`` / ``get_peer_tensor``;
/root/reference/python/triton_dist/utils.py:246-287).  Allocation is a deterministic first-fit over a
Python-side free list: as with NVSHMEM, every rank must perform the same sequence of (de)allocations.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import struct
import tempfile
import uuid
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .. import _C

# ----------------------------------------------------------------------------------------------------------------
# raw pointer -> torch.Tensor through a hand-built DLPack capsule (lets us pick the device for peer mappings)
# ----------------------------------------------------------------------------------------------------------------


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.POINTER(_DLManagedTensor))
_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DELETER)]

_KEEPALIVE = {}


@_DELETER
def _dl_deleter(mt_ptr):
    _KEEPALIVE.pop(C.addressof(mt_ptr.contents), None)


_DL_CODES = {
    torch.float32: (2, 32), torch.float16: (2, 16), torch.float64: (2, 64), torch.bfloat16: (4, 16),
    torch.int8: (0, 8), torch.int16: (0, 16), torch.int32: (0, 32), torch.int64: (0, 64),
    torch.uint8: (1, 8), torch.bool: (6, 8),
}
_PyCapsule_New = C.pythonapi.PyCapsule_New
_PyCapsule_New.restype = C.py_object
_PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_dl_make = None


def _tensor_from_ptr_ctypes(ptr: int, shape: Sequence[int], dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """The original construction (ctypes-built DLManagedTensor, Python deleter): the path every hardware run of this repository
    has exercised; kept for CUDA tensors."""
    view_as = None
    if dtype not in _DL_CODES:   # fp8 & friends: carry as uint8, reinterpret afterwards
        view_as, dtype = dtype, torch.uint8
        assert torch.empty(0, dtype=view_as).element_size() == 1
    shape = [int(s) for s in shape]
    code, bits = _DL_CODES[dtype]
    ndim = len(shape)
    shape_arr = (C.c_int64 * max(ndim, 1))(*shape)
    mt = _DLManagedTensor()
    mt.dl_tensor.data = C.c_void_p(ptr)
    mt.dl_tensor.device = _DLDevice(2 if device.type == "cuda" else 1, device.index or 0)
    mt.dl_tensor.ndim = ndim
    mt.dl_tensor.dtype = _DLDataType(code, bits, 1)
    mt.dl_tensor.shape = C.cast(shape_arr, C.POINTER(C.c_int64))
    mt.dl_tensor.strides = None
    mt.dl_tensor.byte_offset = 0
    mt.manager_ctx = None
    mt.deleter = _dl_deleter
    _KEEPALIVE[C.addressof(mt)] = (mt, shape_arr)
    cap = _PyCapsule_New(C.addressof(mt), b"dltensor", None)
    t = torch.from_dlpack(cap)
    return t.view(view_as) if view_as is not None else t


def tensor_from_ptr(ptr: int, shape: Sequence[int], dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """Alias ``ptr`` as a contiguous tensor of ``shape``/``dtype`` living on ``device`` (no ownership).  Host (emulation backend)
    tensors: the DLPack payload and its deleter live in libtd_host.so (``tdh_dl_make``) -- such a tensor may die on any thread (the DSL
    interpreter runs kernels on Python threads), inside the garbage collector or during interpreter shutdown, where a Python-level
    deleter is not safe to call.  CUDA tensors keep the ctypes construction that all hardware runs have used."""
    global _dl_make
    if device.type == "cuda":
        return _tensor_from_ptr_ctypes(ptr, shape, dtype, device)
    if _dl_make is None:
        lib = _C.host_lib()
        lib.tdh_dl_make.restype = C.c_void_p
        lib.tdh_dl_make.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int]
        _dl_make = lib.tdh_dl_make
    view_as = None
    if dtype not in _DL_CODES:   # fp8 & friends: carry as uint8, reinterpret afterwards
        view_as, dtype = dtype, torch.uint8
        assert torch.empty(0, dtype=view_as).element_size() == 1
    shape = [int(s) for s in shape]
    code, bits = _DL_CODES[dtype]
    ndim = len(shape)
    shape_arr = (C.c_int64 * max(ndim, 1))(*shape)
    mt = _dl_make(C.c_void_p(ptr), ndim, shape_arr, code, bits, 2 if device.type == "cuda" else 1, device.index or 0)
    if not mt:
        raise ValueError(f"tensor_from_ptr: unsupported rank {ndim}")
    cap = _PyCapsule_New(mt, b"dltensor", None)
    t = torch.from_dlpack(cap)
    return t.view(view_as) if view_as is not None else t


# ----------------------------------------------------------------------------------------------------------------
# fd passing between ranks of one node (SCM_RIGHTS)
# ----------------------------------------------------------------------------------------------------------------


def _exchange_fds(my_fd: int, rank: int, world: int, group) -> List[int]:
    """All-to-all exchange of one file descriptor per rank over unix-domain sockets."""
    if world == 1:
        return [my_fd]
    sock_dir = tempfile.gettempdir()
    token = [uuid.uuid4().hex if rank == 0 else None]
    dist.broadcast_object_list(token, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    path = lambda r: os.path.join(sock_dir, f"td_{token[0]}_{r}.sock")
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    if os.path.exists(path(rank)):
        os.unlink(path(rank))
    srv.bind(path(rank))
    srv.listen(world)
    dist.barrier(group=group)
    fds: List[Optional[int]] = [None] * world
    fds[rank] = my_fd
    # deterministic pairing: lower rank connects to higher rank; both directions go over the one connection
    conns = {}
    for peer in range(world):
        if peer > rank:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path(peer))
            c.sendall(struct.pack("i", rank))
            conns[peer] = c
    for _ in range(rank):
        c, _addr = srv.accept()
        (peer,) = struct.unpack("i", c.recv(4, socket.MSG_WAITALL))
        conns[peer] = c
    for peer, c in sorted(conns.items()):
        socket.send_fds(c, [b"f"], [my_fd])
    for peer, c in sorted(conns.items()):
        _msg, got, _flags, _addr = socket.recv_fds(c, 1, 1)
        fds[peer] = got[0]
    dist.barrier(group=group)
    for c in conns.values():
        c.close()
    srv.close()
    try:
        os.unlink(path(rank))
    except OSError:
        pass
    return fds  # type: ignore[return-value]


# ----------------------------------------------------------------------------------------------------------------
# heap
# ----------------------------------------------------------------------------------------------------------------


class SymmetricHeap:
    ALIGN = 1024  # every allocation is 1 KB aligned (TMA / SWIZZLE_128B friendly)

    def __init__(self, nbytes: int, rank: int, world: int, device: torch.device, group=None, multicast: bool = True):
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.is_cuda = device.type == "cuda"
        self.mc_base = 0
        self._handle = None
        if self.is_cuda:
            self._init_cuda(nbytes, multicast)
        else:
            self._init_host(nbytes)
        self._free = [(0, self.nbytes)]   # sorted (offset, size)
        self._live = {}                   # offset -> size
        self._by_ptr = {}                 # data_ptr -> offset

    # ---- backends ----
    def _init_cuda(self, nbytes, multicast):
        lib = _C.cuda_lib()
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        h = lib.td_heap_create(dev, self.rank, self.world, nbytes)
        if not h:
            raise _C.NativeError("td_heap_create: " + lib.td_last_error().decode())
        self._handle = h
        if self.world > 1:
            fd = lib.td_heap_export_fd(h)
            if fd < 0:
                raise _C.NativeError("td_heap_export_fd: " + lib.td_last_error().decode())
            fds = _exchange_fds(fd, self.rank, self.world, self.group)
        else:
            fd, fds = -1, [0]
        arr = (C.c_int * self.world)(*[int(f) for f in fds])
        _C.check(lib.td_heap_map(h, arr), "td_heap_map")
        for r, f in enumerate(fds):
            if self.world > 1 and f is not None and f >= 0:
                os.close(f)
        self.base = int(lib.td_heap_base(h))
        self.stride = int(lib.td_heap_stride(h))
        self.nbytes = int(lib.td_heap_bytes(h))
        if self.world > 1:
            dist.barrier(group=self.group)
        want_mc = multicast and self.world > 1 and os.environ.get("TD_DISABLE_MULTICAST", "0") != "1"
        if want_mc:
            self._init_multicast(lib, dev)

    def _init_multicast(self, lib, dev):
        ok = [bool(lib.td_multicast_supported(dev))]
        oks = [None] * self.world
        dist.all_gather_object(oks, ok[0], group=self.group)
        if not all(oks):
            return
        h = self._handle
        try:
            fd = lib.td_heap_mc_create(h) if self.rank == 0 else -1
            status = [fd >= 0 if self.rank == 0 else True]
            sts = [None] * self.world
            dist.all_gather_object(sts, status[0], group=self.group)
            if not all(sts):
                return
            # ship rank 0's fd to everyone: reuse the all-to-all exchanger (every rank contributes a dummy fd)
            dummy = fd if self.rank == 0 else os.open(os.devnull, os.O_RDONLY)
            fds = _exchange_fds(dummy, self.rank, self.world, self.group)
            # never raise between collectives: a local failure becomes a status every rank sees in the next all-gather,
            # so all ranks take the same (fallback) path instead of pairing mismatched collectives
            imp_rc = lib.td_heap_mc_import(h, fds[0]) if self.rank != 0 else 0
            for f in fds:
                try:
                    os.close(f)
                except OSError:
                    pass
            rc = lib.td_heap_mc_add_device(h) if imp_rc == 0 else 1
            rcs = [None] * self.world
            dist.all_gather_object(rcs, rc, group=self.group)
            if any(rcs):
                return
            rc = lib.td_heap_mc_bind_and_map(h)
            dist.all_gather_object(rcs, rc, group=self.group)
            if any(rcs):
                return
            self.mc_base = int(lib.td_heap_mc_base(h))
        except Exception:
            self.mc_base = 0

    def _init_host(self, nbytes):
        lib = _C.host_lib()
        tok = [uuid.uuid4().hex[:12] if self.rank == 0 else None]
        if self.world > 1:
            dist.broadcast_object_list(tok, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                       group=self.group)
        h = lib.tdh_heap_create(f"tdh_{tok[0]}".encode(), self.rank, self.world, nbytes)
        if not h:
            raise _C.NativeError("tdh_heap_create: " + lib.tdh_last_error().decode())
        self._handle = h
        if self.world > 1:
            dist.barrier(group=self.group)
        if lib.tdh_heap_map(h) != 0:
            raise _C.NativeError("tdh_heap_map: " + lib.tdh_last_error().decode())
        if self.world > 1:
            dist.barrier(group=self.group)
        lib.tdh_heap_unlink(h)
        self.base = int(lib.tdh_heap_base(h))
        self.stride = int(lib.tdh_heap_stride(h))
        self.nbytes = int(lib.tdh_heap_bytes(h))

    # ---- address arithmetic ----
    @property
    def local_base(self) -> int:
        return self.base + self.rank * self.stride

    def offset_of(self, t: torch.Tensor) -> int:
        off = t.data_ptr() - self.local_base
        if not (0 <= off < self.nbytes):
            raise ValueError("tensor does not live in the local symmetric segment")
        return off

    def contains(self, t: torch.Tensor) -> bool:
        off = t.data_ptr() - self.local_base
        return 0 <= off and off + t.numel() * t.element_size() <= self.nbytes

    def peer_ptr(self, ptr_or_tensor, peer: int) -> int:
        ptr = ptr_or_tensor.data_ptr() if isinstance(ptr_or_tensor, torch.Tensor) else int(ptr_or_tensor)
        return ptr + (peer - self.rank) * self.stride

    def mc_ptr(self, ptr_or_tensor) -> int:
        if not self.mc_base:
            raise RuntimeError("multicast (NVLS) mapping not available")
        ptr = ptr_or_tensor.data_ptr() if isinstance(ptr_or_tensor, torch.Tensor) else int(ptr_or_tensor)
        return self.mc_base + (ptr - self.local_base)

    # ---- allocation ----
    def alloc(self, nbytes: int) -> int:
        size = max(self.ALIGN, (int(nbytes) + self.ALIGN - 1) // self.ALIGN * self.ALIGN)
        for i, (off, sz) in enumerate(self._free):
            if sz >= size:
                if sz == size:
                    self._free.pop(i)
                else:
                    self._free[i] = (off + size, sz - size)
                self._live[off] = size
                return off
        raise MemoryError(f"symmetric heap exhausted: need {size} B, free list {self._free[:4]}... "
                          f"(raise TD_SYMM_HEAP_SIZE, currently {self.nbytes} B)")

    def free(self, off: int):
        size = self._live.pop(off)
        self._free.append((off, size))
        self._free.sort()
        merged = []
        for o, s in self._free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self._free = merged

    @property
    def bytes_in_use(self) -> int:
        return sum(self._live.values())

    def tensor(self, shape, dtype: torch.dtype, zero: bool = True) -> torch.Tensor:
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
        numel = 1
        for s in shape:
            numel *= s
        nbytes = numel * torch.empty(0, dtype=dtype).element_size()
        off = self.alloc(nbytes)
        t = tensor_from_ptr(self.local_base + off, shape, dtype, self.device)
        self._by_ptr[t.data_ptr()] = off
        if zero and numel:
            t.zero_()
        return t

    def peer_view(self, t: torch.Tensor, peer: int) -> torch.Tensor:
        """The tensor at the same heap offset in ``peer``'s segment, addressable from this process."""
        if peer == self.rank:
            return t
        assert t.is_contiguous()
        return tensor_from_ptr(self.peer_ptr(t, peer), t.shape, t.dtype, self.device)

    def mc_view(self, t: torch.Tensor) -> torch.Tensor:
        return tensor_from_ptr(self.mc_ptr(t), t.shape, t.dtype, self.device)

    def free_tensor(self, t: torch.Tensor):
        off = self._by_ptr.pop(t.data_ptr(), None)
        if off is None:
            raise ValueError("not a symmetric tensor returned by this heap")
        self.free(off)

    def destroy(self):
        if self._handle is None:
            return
        if self.is_cuda:
            torch.cuda.synchronize()
            _C.cuda_lib().td_heap_destroy(self._handle)
        else:
            _C.host_lib().tdh_heap_destroy(self._handle)
        self._handle = None
