"""CPU interpreter for DSL kernels: the Python function itself is the semantics.

Every CUDA thread of a block is a Python thread (``__syncthreads`` is a ``threading.Barrier``, a warp shuffle is an exchange buffer
per warp), the blocks of a small grid run concurrently (larger grids block by block), pointers are views over flat CPU tensors, shared arrays are per-block numpy arrays.  The
distributed primitives (``symm_at`` / ``notify`` / ``wait``) map onto the emulation backend of this framework
(``triton_dist.language`` over the shared-memory heap), so a multi-rank DSL kernel can be executed by N CPU processes -- the same way
``tests/test_dist_cpu.py`` runs every fused-op protocol.  TMA / tcgen05 intrinsics have no CPU meaning and raise.

The reference has no counterpart (its DSL kernels only run on a GPU, python/little_kernel/tests/conftest.py:50-63 gates them); this is
what lets the DSL be tested in the CPU-only tier.
"""
from __future__ import annotations

import threading
from typing import Any, Dict, List, Tuple

import numpy as np

from . import types as T

_tls = threading.local()
_atomic_lock = threading.Lock()


class BlockState:
    def __init__(self, nthreads: int):
        self.nthreads = nthreads
        self.barrier = threading.Barrier(nthreads)
        self.shared: Dict[Any, np.ndarray] = {}
        self.lock = threading.Lock()
        nwarps = (nthreads + 31) // 32
        self.warp_size = [min(32, nthreads - w * 32) for w in range(nwarps)]
        self.warp_barrier = [threading.Barrier(n) for n in self.warp_size]
        self.warp_xchg: List[List[Any]] = [[None] * 32 for _ in range(nwarps)]
        self.error = None
        # Blackwell async-pipeline model (see ``pipeline.py``): mbarriers keyed by (shared-array key, index), a 128-lane x 512-column
        # tensor memory, and the CTAs of this block's cluster
        self.mbarriers: Dict[Any, Any] = {}
        self.tmem = np.zeros((128, 512), dtype=np.float32)
        self.cluster: List["BlockState"] = [self]
        self.cta_rank = 0
        self.cluster_barrier = self.barrier


class ThreadCtx:
    def __init__(self, tid, bid, bdim, gdim, block: BlockState, dyn_smem: np.ndarray):
        self.tid, self.bid, self.bdim, self.gdim, self.block = tid, bid, bdim, gdim, block
        self.linear = tid[0] + bdim[0] * (tid[1] + bdim[1] * tid[2])
        self.dyn_smem = dyn_smem
        self.dyn_off = 0


def cur() -> ThreadCtx:
    c = getattr(_tls, "ctx", None)
    if c is None:
        raise RuntimeError("DSL intrinsic called outside a kernel (use kernel.interpret(...) on CPU or launch it on a GPU)")
    return c


def active() -> bool:
    return getattr(_tls, "ctx", None) is not None


# ------------------------------------------------------------------------------------------------------------
class Ptr:
    """A typed pointer into a flat CPU tensor (or numpy array): ``p[i]``, ``p[i] = v``, ``p + n``."""

    def __init__(self, base, off: int = 0, elem: T.Type = None, owner=None):
        self.base, self.off, self.elem, self.owner = base, int(off), elem, owner if owner is not None else base

    def __getitem__(self, i):
        v = self.base[self.off + int(i)]
        v = v.item() if hasattr(v, "item") else v
        e = self.elem
        if e is not None and getattr(e, "kind", None) == "u" and isinstance(v, int) and v < 0:
            v += 1 << e.bits                          # an unsigned pointer over signed storage (torch has no uint32 / uint64 arithmetic)
        return v

    def __setitem__(self, i, v):
        e = self.elem
        if e is not None and getattr(e, "kind", None) == "u" and isinstance(v, int) and hasattr(self.base, "dtype") \
                and getattr(self.base.dtype, "is_signed", False) and v >= 1 << (e.bits - 1):
            v -= 1 << e.bits                          # same bits, representable in the signed storage
        self.base[self.off + int(i)] = v

    skey = None                                      # set for pointers into shared arrays (mbarrier identity)

    def __add__(self, n):
        p = Ptr(self.base, self.off + int(n), self.elem, self.owner)
        p.skey = self.skey
        return p

    __radd__ = __add__

    def __sub__(self, n):
        if isinstance(n, Ptr):
            return self.off - n.off
        return Ptr(self.base, self.off - int(n), self.elem, self.owner)

    def reinterpret(self, elem: T.Type):
        """``reinterpret_cast<elem*>(p)``: the same bytes viewed as another element type (torch bases; byte offset must be aligned)."""
        import torch
        b = self.base
        if not isinstance(b, torch.Tensor) or getattr(elem, "torch_name", None) is None:
            return self
        new_dt = getattr(torch, elem.torch_name)
        if new_dt == b.dtype:
            return self
        raw = b[self.off:].view(torch.uint8)
        es = torch.empty(0, dtype=new_dt).element_size()
        p = Ptr(raw[: raw.numel() // es * es].view(new_dt), 0, elem, self.owner)
        p.skey = self.skey
        return p

    def tensor(self):
        """The underlying torch tensor from this element on (for the host mirror of the distributed primitives)."""
        return self.base[self.off:]

    def _addr(self):
        b = self.base
        a = b.data_ptr() if hasattr(b, "data_ptr") else b.__array_interface__["data"][0]
        es = b.element_size() if hasattr(b, "element_size") else b.itemsize
        return a + self.off * es

    def __eq__(self, o):
        return isinstance(o, Ptr) and self._addr() == o._addr()

    def __ne__(self, o):
        return not self.__eq__(o)

    def __lt__(self, o):
        return self._addr() < o._addr()

    __hash__ = None


class SharedArray:
    """A (view of a) per-block shared array.  ``key`` names the allocation -- identical in every CTA of a launch, which is how the
    cluster model finds "the same address in the peer CTA" -- and ``base`` is the element offset of this view inside it."""

    def __init__(self, arr: np.ndarray, shape, key=None, base: int = 0):
        self.arr, self.shape, self.key, self.base = arr, tuple(shape), key, int(base)

    def _flat(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        f = 0
        for k, i in enumerate(idx):
            stride = 1
            for d in self.shape[k + 1:]:
                stride *= d
            f += int(i) * stride
        return f

    def __getitem__(self, idx):
        n = len(idx) if isinstance(idx, tuple) else 1
        if n < len(self.shape):                      # fewer indices than dimensions: a row view (``sA[stage]``)
            return SharedArray(self.arr, self.shape[n:], self.key, self.base + self._flat(idx))
        v = self.arr[self.base + self._flat(idx)]
        return v.item() if hasattr(v, "item") else v

    def __setitem__(self, idx, v):
        self.arr[self.base + self._flat(idx)] = v

    def __add__(self, n):
        p = Ptr(self.arr, self.base + int(n))
        p.skey = self.key
        return p

    @property
    def numel(self):
        n = 1
        for d in self.shape:
            n *= d
        return n


_NP = {"bool": np.bool_, "i8": np.int8, "u8": np.uint8, "i16": np.int16, "u16": np.uint16, "i32": np.int32, "u32": np.uint32,
       "i64": np.int64, "u64": np.uint64, "f16": np.float16, "bf16": np.float32, "f32": np.float32, "f64": np.float64, "e4m3": np.float32}


def np_dtype(t: T.Type):
    return _NP.get(getattr(t, "name", ""), np.uint8)


def shared_array(key, shape, dtype, per_block=True) -> SharedArray:
    c = cur()
    shape = T.shape_tuple(shape)
    n = int(np.prod(shape))
    if not per_block:
        return SharedArray(np.zeros(n, dtype=np_dtype(dtype)), shape)
    with c.block.lock:
        arr = c.block.shared.get(key)
        if arr is None:
            arr = c.block.shared[key] = np.zeros(n, dtype=np_dtype(dtype))
    return SharedArray(arr, shape, key)


def syncthreads():
    cur().block.barrier.wait()


def warp_exchange(value, src_lane_of):
    """All lanes of the calling warp publish ``value``; lane l gets the value of lane ``src_lane_of(l)`` (its own when out of range)."""
    c = cur()
    w, lane = c.linear // 32, c.linear % 32
    blk = c.block
    blk.warp_xchg[w][lane] = value
    blk.warp_barrier[w].wait()
    src = src_lane_of(lane)
    out = blk.warp_xchg[w][src] if 0 <= src < blk.warp_size[w] else value
    blk.warp_barrier[w].wait()
    return out


def warp_collect(value) -> list:
    c = cur()
    w, lane = c.linear // 32, c.linear % 32
    blk = c.block
    blk.warp_xchg[w][lane] = value
    blk.warp_barrier[w].wait()
    out = list(blk.warp_xchg[w][:blk.warp_size[w]])
    blk.warp_barrier[w].wait()
    return out


def atomic_rmw(p: Ptr, i, fn):
    """``old = p[i]; p[i] = fn(old); return old`` atomically.  32-bit integers in torch tensors go through a hardware CAS loop
    (libtd_host): the word may live on the emulation backend's shared-memory heap, where the other parties are other PROCESSES and a
    Python lock would protect nothing.  Everything else (floats, 64-bit, numpy shared arrays) is only ever contended by the threads
    of this interpreter and takes the lock."""
    base = p.base
    if hasattr(base, "data_ptr") and base.element_size() == 4 and not base.dtype.is_floating_point:
        from .. import _C
        lib = _C.host_lib()
        addr = base.data_ptr() + (p.off + int(i)) * 4
        signed = base.dtype.is_signed
        while True:
            raw = int(lib.tdh_ld_acquire32(addr))
            old = raw - (1 << 32) if signed and raw >> 31 else raw
            if int(lib.tdh_atomic_cas32(addr, raw, int(fn(old)) & 0xFFFFFFFF)) == raw:
                return old
    with _atomic_lock:
        old = p[i]
        p[i] = fn(old)
    return old


# ------------------------------------------------------------------------------------------------------------
def _wrap_arg(a, ty):
    import torch
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise ValueError("interpret() runs on CPU tensors")
        if a.is_contiguous():
            flat = a.view(-1)
        else:       # a strided view (say, the k heads of a packed qkv tensor): point at its first element inside the WHOLE storage, as the
            # device pointer would -- the kernel brings its own strides, and its stores must land in the caller's tensor
            n = a.untyped_storage().nbytes() // a.element_size() - a.storage_offset()
            flat = torch.as_strided(a, (n,), (1,))
        return Ptr(flat, 0, getattr(ty, "elem", None), owner=a)
    if isinstance(ty, T.Scalar):
        return ty.wrap(a)
    return a


MAX_CONCURRENT_THREADS = 2048


def run(fn, param_types, grid, block, args, dyn_smem_bytes: int = 0, cluster=(1, 1, 1)):
    """Blocks of a small grid (<= MAX_CONCURRENT_THREADS emulated threads in total) run concurrently, so kernels whose blocks depend on
    each other (or on other ranks, block by block) make progress as on a GPU; larger grids run one block after the other."""
    grid = tuple(grid) + (1,) * (3 - len(tuple(grid))) if not isinstance(grid, int) else (grid, 1, 1)
    block = tuple(block) + (1,) * (3 - len(tuple(block))) if not isinstance(block, int) else (block, 1, 1)
    nthreads = block[0] * block[1] * block[2]
    wrapped = [_wrap_arg(a, t) for a, t in zip(args, param_types)]
    bids = [(bx, by, bz) for bz in range(grid[2]) for by in range(grid[1]) for bx in range(grid[0])]
    tids: List[Tuple[int, int, int]] = [(x, y, z) for z in range(block[2]) for y in range(block[1]) for x in range(block[0])]
    errors: List[BaseException] = []

    def body(tid, bid, blk, dyn):
        _tls.ctx = ThreadCtx(tid, bid, block, grid, blk, dyn)
        try:
            fn(*wrapped)
        except threading.BrokenBarrierError:
            pass
        except BaseException as e:      # noqa: BLE001
            errors.append(e)
            for member in blk.cluster:
                member.barrier.abort()
                member.cluster_barrier.abort()
                for b in member.warp_barrier:
                    b.abort()
                for mb in list(member.mbarriers.values()):
                    mb.abort()
        finally:
            _tls.ctx = None

    csize = int(cluster[0]) * int(cluster[1]) * int(cluster[2]) if not isinstance(cluster, int) else int(cluster)
    if csize > 1 and (cluster[1] != 1 or cluster[2] != 1 or grid[0] % csize):
        raise ValueError("the interpreter models clusters along x only (grid.x must be a multiple of the cluster size)")
    per_wave = max(1, MAX_CONCURRENT_THREADS // nthreads) if len(bids) * nthreads <= MAX_CONCURRENT_THREADS else 1
    per_wave = max(csize, per_wave // csize * csize)              # the CTAs of a cluster always run together
    for w0 in range(0, len(bids), per_wave):
        ths = []
        wave = bids[w0:w0 + per_wave]
        blocks = [BlockState(nthreads) for _ in wave]
        if csize > 1:
            for c0 in range(0, len(blocks), csize):
                members = blocks[c0:c0 + csize]
                cb = threading.Barrier(nthreads * len(members))
                for r, b in enumerate(members):
                    b.cluster, b.cta_rank, b.cluster_barrier = members, r, cb
        for bid, blk in zip(wave, blocks):
            dyn = np.zeros(max(dyn_smem_bytes, 1), dtype=np.uint8)
            if nthreads == 1 and per_wave == 1:
                body(tids[0], bid, blk, dyn)
                continue
            ths += [threading.Thread(target=body, args=(t, bid, blk, dyn), daemon=True) for t in tids]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errors:
            raise errors[0]
