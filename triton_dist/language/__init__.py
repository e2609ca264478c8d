"""Host-side mirror of the device language.

In the reference these are Triton builtins lowered through an MLIR dialect
(/root/reference/python/triton_dist/language/distributed_ops.py:53-107).  In this framework the real primitives
are C++ ``__device__`` functions (csrc/td/primitives.cuh: ``td::wait / notify / symm_at / rank / num_ranks``) used
inside the hand-written kernels.  This module offers the same vocabulary at Python level -- stream-ordered one-thread
kernels on the GPU, atomics on the shared-memory heap without one -- for tutorials, tests and host-driven protocols
(pipeline-parallel hand-off, copy-engine producers).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _C
from .. import utils as U

_C.register("td_signal", C.c_int, [C.c_void_p, C.c_uint, C.c_int, C.c_void_p])
_C.register("td_wait", C.c_int, [C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_void_p])

SIGNAL_OP = {"set": 1, "add": 2}
COMM_SCOPE = {"gpu": 1, "intra_node": 2, "inter_node": 3}


def rank(axis: int = -1) -> int:
    return U.rank()


def num_ranks(axis: int = -1) -> int:
    return U.world_size()


def symm_at(t: torch.Tensor, peer: int) -> torch.Tensor:
    return U.symm_at(t, peer)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def notify(ptr: torch.Tensor, peer: int, signal: int = 1, sig_op: str = "set", comm_scope: str = "intra_node"):
    """Write/add ``signal`` into element 0 of ``ptr`` on rank ``peer`` with release semantics: everything this
    stream (GPU) / thread (host) wrote before is visible to whoever acquires the flag."""
    if comm_scope == "inter_node":
        raise NotImplementedError("single NVSwitch domain only (SURVEY.md 5.8)")
    heap = U.get_heap()
    addr = heap.peer_ptr(ptr, peer) if peer != heap.rank else ptr.data_ptr()
    if ptr.is_cuda:
        _C.check(_C.cuda_lib().td_signal(C.c_void_p(addr), int(signal) & 0xFFFFFFFF, SIGNAL_OP[sig_op], _stream()), "td_signal")
    else:
        _C.host_lib().tdh_notify32(C.c_void_p(addr), int(signal) & 0xFFFFFFFF, SIGNAL_OP[sig_op])


def wait(barrier_ptrs: torch.Tensor, num_barriers: int = 1, scope: str = "sys", semantic: str = "acquire",
         wait_value: int = 1, geq: bool = False) -> int:
    """Block (the stream / the thread) until ``num_barriers`` consecutive flags equal ``wait_value``.  Returns a
    token for :func:`consume_token`, as in the reference."""
    if barrier_ptrs.is_cuda:
        _C.check(_C.cuda_lib().td_wait(C.c_void_p(barrier_ptrs.data_ptr()), num_barriers, int(wait_value) & 0xFFFFFFFF,
                                       int(geq), _stream()), "td_wait")
    else:
        rc = _C.host_lib().tdh_wait32_n(C.c_void_p(barrier_ptrs.data_ptr()), num_barriers, int(wait_value) & 0xFFFFFFFF,
                                        1 if geq else 0, U.get_int_env("TD_HOST_TIMEOUT_US", 60_000_000))
        if rc:
            raise TimeoutError("dl.wait timed out (hang detected)")
    return int(wait_value)


def consume_token(value, token):
    """Identity; keeps the data dependence explicit in protocol code (DistributedOpToLLVM.cpp:231-241)."""
    return value


import contextlib


@contextlib.contextmanager
def simt_exec_region():
    """Device-side concept of the reference DSL (a region executed per thread instead of per block).  The CUDA kernels
    of this framework are written in SIMT form already, so the host mirror is a no-op context."""
    yield


# ---- per-thread vectors of the SIMT region (reference: language/simt_ops.py:35-286) --------------------------------------------------
class vector:
    """A short fixed-length vector held by one thread: element-wise ``+ - *``, ``to`` (convert), ``recast`` (reinterpret the bits).
    In device code of this framework that is a register array (``ll.local`` in the DSL, a C array in CUDA); this host mirror backs the
    emulation tests and documentation examples."""

    def __init__(self, data):
        self.data = data if isinstance(data, torch.Tensor) else torch.as_tensor(data)

    def _bin(self, other, op):
        o = other.data if isinstance(other, vector) else other
        return vector(op(self.data, o))

    def __add__(self, o):
        return self._bin(o, torch.add)

    def __sub__(self, o):
        return self._bin(o, torch.sub)

    def __mul__(self, o):
        return self._bin(o, torch.mul)

    __radd__, __rmul__ = __add__, __mul__

    def __getitem__(self, i):
        return self.data[i]

    def __setitem__(self, i, v):
        self.data[i] = v

    def __len__(self):
        return self.data.numel()

    def to(self, dtype):
        return vector(self.data.to(dtype))

    def recast(self, dtype):
        return vector(self.data.contiguous().view(dtype))

    @property
    def dtype(self):
        return self.data.dtype


def make_vector(values, dtype=None) -> vector:
    return vector(torch.as_tensor(list(values), dtype=dtype))


def zeros_vector(n: int, dtype=torch.float32) -> vector:
    return vector(torch.zeros(n, dtype=dtype))


def extern_call(lib, symbol: str, args=(), restype=None, argtypes=None):
    """Call ``symbol`` of a shared library (path or loaded ``ctypes.CDLL``) -- the host counterpart of the reference's device-side
    ``extern_call`` (language/core.py:85-116); inside DSL kernels use ``triton_dist.lk.stdlib.extern_call``."""
    import ctypes
    handle = ctypes.CDLL(lib) if isinstance(lib, str) else lib
    fn = getattr(handle, symbol)
    fn.restype = restype
    if argtypes is not None:
        fn.argtypes = list(argtypes)
    return fn(*args)

