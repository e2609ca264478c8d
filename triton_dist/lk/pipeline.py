"""A functional model of the Blackwell asynchronous pipeline for the CPU interpreter: mbarriers (arrival counts, transaction bytes,
phase parity), TMA tile loads through tensor maps (incl. ``cta_group::2`` loads that signal the leader CTA), tensor memory, single-thread
``tcgen05.mma`` (``cta_group::1`` and ``::2``), ``tcgen05.commit`` (incl. the 2-CTA multicast) and ``tcgen05.ld``.

Purpose: run the DSL's tcgen05 GEMM ladder on CPU tensors with a handful of tiles and check what cannot be seen without hardware --
barrier phases, stage indexing across tile boundaries, accumulator double buffering, the cluster protocol -- against ``A @ B^T``.  A
protocol error shows up as wrong numbers or as a reported deadlock (every wait has a timeout), not as a hung GPU.

What is modelled is the *logical* behaviour, not the byte layout: a staged tile is a row-major ``[rows, 64]`` array (the 128-byte
swizzle is a layout detail between TMA and the MMA that cancels out), a shared-memory descriptor is (allocation, element offset), and
advancing a descriptor by 2 (32 bytes) moves 16 elements along K, as the kernels do.  MMAs execute at issue, so ``commit`` arrives at
once; TMA loads complete at issue.  Operations keep program order per thread, which is a legal schedule of the real machine -- the model
cannot find races that depend on asynchronous completion orders, it finds the protocol and indexing errors.
"""
from __future__ import annotations

import os
import struct
import threading
from typing import Optional

import numpy as np

from . import interp as I

TIMEOUT_S = float(os.environ.get("TD_LK_INTERP_TIMEOUT_S", "120"))
K_TILE = 64                         # elements per staged tile row (128 bytes of bf16: one SWIZZLE_128B atom)


class Deadlock(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------------------
# mbarrier
# ------------------------------------------------------------------------------------------------------------
class MBarrier:
    def __init__(self, count: int, name):
        self.count, self.name = int(count), name
        self.pending, self.tx, self.phase = int(count), 0, 0          # phase = number of completed phases
        self.cv = threading.Condition()
        self.dead = False

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count
            self.cv.notify_all()

    def arrive(self, n: int = 1, tx: int = 0):
        with self.cv:
            self.tx += tx
            self.pending -= n
            if self.pending < 0:
                raise RuntimeError(f"mbarrier {self.name}: more arrivals than its count ({self.count}) in one phase")
            self._maybe_complete()

    def expect_tx(self, nbytes: int):
        with self.cv:
            self.tx += nbytes

    def complete_tx(self, nbytes: int):
        with self.cv:
            self.tx -= nbytes
            self._maybe_complete()

    def test(self, parity: int) -> bool:
        return (self.phase & 1) != (int(parity) & 1)

    def wait(self, parity: int):
        with self.cv:
            if not self.cv.wait_for(lambda: self.test(parity) or self.dead, timeout=TIMEOUT_S):
                c = I.cur()
                raise Deadlock(f"mbarrier {self.name} (count {self.count}): block {c.bid} thread {c.linear} waited {TIMEOUT_S:.0f} s for "
                               f"parity {parity}; phase {self.phase}, pending arrivals {self.pending}, pending tx bytes {self.tx}")
            if self.dead and not self.test(parity):
                raise threading.BrokenBarrierError

    def abort(self):
        with self.cv:
            self.dead = True
            self.cv.notify_all()


def _bar_key(p):
    if isinstance(p, I.SharedArray) and p.key is not None:
        return (p.key, p.base)
    if not isinstance(p, I.Ptr) or p.skey is None:
        raise TypeError("mbarrier operations take a pointer into a shared array (``bars + i``)")
    return (p.skey, p.off)


def _bar_of(block, key, create_count: Optional[int] = None) -> MBarrier:
    with block.lock:
        b = block.mbarriers.get(key)
        if b is None or create_count is not None:
            if create_count is None:
                raise RuntimeError(f"mbarrier {key} used before mbar_init")
            b = block.mbarriers[key] = MBarrier(create_count, key)
    return b


def _wait_for_init(block, key) -> MBarrier:
    """Another CTA of the cluster may touch my barrier only after a cluster_sync that follows the init, so it exists; be defensive."""
    return _bar_of(block, key)


def mbar_init(p, count):
    _bar_of(I.cur().block, _bar_key(p), int(count))


def mbar_arrive(p):
    _bar_of(I.cur().block, _bar_key(p)).arrive()


def mbar_arrive_cluster(p, cta):
    blk = I.cur().block
    _wait_for_init(blk.cluster[int(cta)], _bar_key(p)).arrive()


def mbar_arrive_expect_tx(p, nbytes):
    _bar_of(I.cur().block, _bar_key(p)).arrive(1, int(nbytes))


def mbar_expect_tx(p, nbytes):
    _bar_of(I.cur().block, _bar_key(p)).expect_tx(int(nbytes))


def mbar_wait(p, parity):
    _bar_of(I.cur().block, _bar_key(p)).wait(int(parity))


def mbar_try_wait(p, parity):
    return _bar_of(I.cur().block, _bar_key(p)).test(int(parity))


def cluster_sync():
    I.cur().block.cluster_barrier.wait()


# ------------------------------------------------------------------------------------------------------------
# tensor maps / TMA
# ------------------------------------------------------------------------------------------------------------
class HostTensorMap:
    """What ``lk.tma_2d`` returns for a CPU tensor: a 2-D row-major tensor + the box (inner = contiguous elements, outer = rows)."""

    def __init__(self, t, box_inner: int, box_outer: int, swizzle: int = 128):
        assert t.dim() == 2 and t.stride(1) == 1
        self.t, self.box_inner, self.box_outer = t, int(box_inner), int(box_outer)
        self.nbytes = self.box_inner * self.box_outer * t.element_size()


def _smem_view(dst) -> I.SharedArray:
    if isinstance(dst, I.SharedArray):
        return dst
    raise TypeError("TMA destination / MMA operand must be a shared array (row view)")


def _tma_fill(tmap: HostTensorMap, dst, c_inner: int, c_outer: int):
    view = _smem_view(dst)
    rows, cols = tmap.box_outer, tmap.box_inner
    tile = np.zeros((rows, cols), dtype=np.float32)
    t = tmap.t
    r0, r1 = max(0, int(c_outer)), min(t.shape[0], int(c_outer) + rows)
    k0, k1 = max(0, int(c_inner)), min(t.shape[1], int(c_inner) + cols)
    if r1 > r0 and k1 > k0:                                            # out-of-bounds elements are zero-filled
        tile[r0 - int(c_outer):r1 - int(c_outer), k0 - int(c_inner):k1 - int(c_inner)] = t[r0:r1, k0:k1].float().numpy()
    n = rows * cols
    if view.numel < n:
        raise RuntimeError(f"TMA box ({rows} x {cols}) does not fit the shared-memory tile ({view.numel} elements)")
    view.arr[view.base:view.base + n] = tile.reshape(-1)


def tma_load_2d(tmap, bar, dst, c_inner, c_outer, *_):
    _tma_fill(tmap, dst, c_inner, c_outer)
    _bar_of(I.cur().block, _bar_key(bar)).complete_tx(tmap.nbytes)


def tma_load_2d_2sm(tmap, bar, dst, c_inner, c_outer, *_):
    """cta_group::2 load: the bytes land in MY shared memory, the transaction completes on the LEADER CTA's barrier."""
    _tma_fill(tmap, dst, c_inner, c_outer)
    blk = I.cur().block
    _wait_for_init(blk.cluster[0], _bar_key(bar)).complete_tx(tmap.nbytes)


# ------------------------------------------------------------------------------------------------------------
# shared-memory descriptors, tensor memory, tcgen05
# ------------------------------------------------------------------------------------------------------------
class SmemAddr:
    def __init__(self, key, off):
        self.key, self.off = key, int(off)


class SmemDesc:
    """K-major, 128-byte-swizzled operand tile starting at (allocation ``key``, element ``off``); ``+ n`` advances n * 16 bytes along K."""

    def __init__(self, key, off, k_off: int = 0):
        self.key, self.off, self.k_off = key, int(off), int(k_off)

    def __add__(self, n):
        return SmemDesc(self.key, self.off, self.k_off + int(n) * 8)

    __radd__ = __add__

    def rows(self, block, nrows: int, k_elems: int) -> np.ndarray:
        arr = block.shared[self.key]
        tile = arr[self.off:self.off + nrows * K_TILE].reshape(nrows, K_TILE)
        if self.k_off + k_elems > K_TILE:
            raise RuntimeError("descriptor advanced past the 64-element swizzle atom")
        return tile[:, self.k_off:self.k_off + k_elems]


def smem_addr(x) -> SmemAddr:
    v = _smem_view(x)
    return SmemAddr(v.key, v.base)


def make_smem_desc_k128(a: SmemAddr) -> SmemDesc:
    return SmemDesc(a.key, a.off)


def tmem_alloc(slot, ncols, cta_group=1):
    c = I.cur()
    if c.linear % 32 == 0:
        slot[0] = 0                                   # the model gives every CTA the whole 512-column tensor memory at address 0
    assert int(ncols) <= 512 and (int(ncols) & (int(ncols) - 1)) == 0 and int(ncols) >= 32, "TMEM columns: power of two in [32, 512]"


def _idesc_mn(idesc: int):
    return ((int(idesc) >> 24) & 0x1F) * 16, ((int(idesc) >> 17) & 0x3F) * 8


def mma_f16(d_tmem, adesc: SmemDesc, bdesc: SmemDesc, idesc, accumulate, cta_group=1):
    """D[M, N] (+)= A[M, 16] B[N, 16]^T into tensor memory.  cta_group 2: rows 0..127 of A and the first N/2 rows of B come from the
    leader's shared memory, the rest from the peer CTA at the same offsets; D rows 128..255 live in the peer's tensor memory."""
    blk = I.cur().block
    M, N = _idesc_mn(idesc)
    col = int(d_tmem) & 0xFFFF
    ctas = blk.cluster[:2] if int(cta_group) == 2 else [blk]
    if int(cta_group) == 2 and blk.cta_rank != 0:
        raise RuntimeError("tcgen05.mma.cta_group::2 must be issued by the leader CTA of the pair")
    mp, nb = M // len(ctas), N // len(ctas)
    if mp != 128 and not (len(ctas) == 1 and mp == 64):
        raise RuntimeError(f"unsupported MMA shape M = {M} for cta_group {cta_group}")
    B = np.concatenate([bdesc.rows(c, nb, 16) for c in ctas], axis=0)                 # [N, 16]
    for c in ctas:
        A = adesc.rows(c, mp, 16)                                                     # [128, 16]
        d = A @ B.T
        if int(accumulate):
            c.tmem[:mp, col:col + N] += d
        else:
            c.tmem[:mp, col:col + N] = d


def mma_commit(bar):
    _bar_of(I.cur().block, _bar_key(bar)).arrive()


def mma_commit_2sm(bar, mask):
    blk = I.cur().block
    for r, c in enumerate(blk.cluster[:2]):
        if (int(mask) >> r) & 1:
            _wait_for_init(c, _bar_key(bar)).arrive()


def tmem_ld_32x32b_x32(taddr, regs, n: int = 32):
    c = I.cur()
    lane = c.linear % 32
    row, col = (int(taddr) >> 16) + lane, int(taddr) & 0xFFFF
    if (int(taddr) >> 16) != (c.linear // 32 % 4) * 32:
        raise RuntimeError(f"warp {c.linear // 32} may only read TMEM lanes {(c.linear // 32 % 4) * 32}..+31, asked for {int(taddr) >> 16}")
    vals = c.block.tmem[row, col:col + n]
    for i in range(n):
        regs[i] = struct.unpack("I", struct.pack("f", float(vals[i])))[0]


# ------------------------------------------------------------------------------------------------------------
# the epilogue's packing / vector stores
# ------------------------------------------------------------------------------------------------------------
def _bf16_bits(f: float) -> int:
    b = struct.unpack("I", struct.pack("f", float(f)))[0]
    if (b & 0x7F800000) == 0x7F800000:               # inf / nan
        return (b >> 16) & 0xFFFF
    return ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) & 0xFFFF


def pack_bf16x2(lo, hi) -> int:
    return _bf16_bits(lo) | (_bf16_bits(hi) << 16)


def _bf16_val(bits: int) -> float:
    return struct.unpack("f", struct.pack("I", (int(bits) & 0xFFFF) << 16))[0]


def _torch_bytes(p):
    """(uint8 view of the tensor behind ``p``, byte offset) for pointers into torch tensors of any element type; None otherwise."""
    base = getattr(p, "base", None)
    if base is None or not hasattr(base, "data_ptr"):
        return None
    import torch
    return base.view(torch.uint8), p.off * base.element_size()


def st_v4(dst, v):
    """16-byte store of the four packed 32-bit words of ``v`` -- byte exact, whatever the element type of the tensor ``dst`` points into."""
    tb = _torch_bytes(dst)
    if tb is not None:
        import torch
        raw, off = tb
        words = torch.tensor([int(v.x) & 0xFFFFFFFF, int(v.y) & 0xFFFFFFFF, int(v.z) & 0xFFFFFFFF, int(v.w) & 0xFFFFFFFF], dtype=torch.int64)
        raw[off:off + 16] = words.to(torch.int32).view(torch.uint8)
        return
    vals = []                                            # numpy-backed (shared) arrays of 16-bit floats
    for w in (v.x, v.y, v.z, v.w):
        vals += [_bf16_val(w), _bf16_val(int(w) >> 16)]
    for i, x in enumerate(vals):
        dst[i] = x


def ld_v4(src):
    """16-byte load -> the four packed 32-bit words (byte exact for pointers into torch tensors of any element type)."""
    import types
    tb = _torch_bytes(src)
    if tb is not None:
        import torch
        raw, off = tb
        w = raw[off:off + 16].clone().view(torch.int32).tolist()
        return types.SimpleNamespace(x=w[0] & 0xFFFFFFFF, y=w[1] & 0xFFFFFFFF, z=w[2] & 0xFFFFFFFF, w=w[3] & 0xFFFFFFFF)
    f = [src[i] for i in range(8)]
    return types.SimpleNamespace(x=pack_bf16x2(f[0], f[1]), y=pack_bf16x2(f[2], f[3]), z=pack_bf16x2(f[4], f[5]), w=pack_bf16x2(f[6], f[7]))


def red_add_bf16x8(dst, v):
    """``red.global.add.noftz.v4.bf16x2``: eight bf16 additions, each 32-bit word updated atomically (CAS loop on the word through the
    host library, so concurrent adders in OTHER PROCESSES -- ranks of the emulation backend -- are handled like on the GPU)."""
    import ctypes
    from .. import _C
    lib = _C.host_lib()
    base = dst.base
    es = base.element_size()
    assert es == 2 and dst.off % 2 == 0, "red_add_bf16x8 needs a 4-byte aligned pointer into a 16-bit tensor"
    addr = base.data_ptr() + dst.off * es
    for i, w in enumerate((v.x, v.y, v.z, v.w)):
        a = ctypes.c_void_p(addr + 4 * i)
        while True:
            old = int(lib.tdh_ld_acquire32(a))
            lo = _bf16_bits(_bf16_val(old) + _bf16_val(w))
            hi = _bf16_bits(_bf16_val(old >> 16) + _bf16_val(int(w) >> 16))
            if int(lib.tdh_atomic_cas32(a, old, lo | (hi << 16))) == old:
                break


def mma_m16n8k16_bf16(acc, off, a0, a1, a2, a3, b0, b1):
    """Warp-collective ``mma.sync.m16n8k16`` (bf16 x bf16 -> fp32) following the PTX fragment layout: every lane contributes its A / B
    registers (packed bf16 pairs); lane 4g + tig receives D(g, 2tig..2tig+1) and D(g + 8, ...) accumulated into ``acc[off .. off+3]``."""
    regs = I.warp_collect((int(a0), int(a1), int(a2), int(a3), int(b0), int(b1)))
    if len(regs) != 32:
        raise RuntimeError("mma.sync needs a full warp")
    A = np.zeros((16, 16), np.float32)
    B = np.zeros((16, 8), np.float32)
    for lane, (r0, r1, r2, r3, q0, q1) in enumerate(regs):
        g, tig = lane >> 2, lane & 3
        for w, (row, col) in ((r0, (g, 2 * tig)), (r1, (g + 8, 2 * tig)), (r2, (g, 2 * tig + 8)), (r3, (g + 8, 2 * tig + 8))):
            A[row, col], A[row, col + 1] = _bf16_val(w), _bf16_val(w >> 16)
        for w, k in ((q0, 2 * tig), (q1, 2 * tig + 8)):
            B[k, g], B[k + 1, g] = _bf16_val(w), _bf16_val(w >> 16)
    D = A @ B
    lane = I.cur().linear % 32
    g, tig = lane >> 2, lane & 3
    off = int(off)
    acc[off + 0] = acc[off + 0] + float(D[g, 2 * tig])
    acc[off + 1] = acc[off + 1] + float(D[g, 2 * tig + 1])
    acc[off + 2] = acc[off + 2] + float(D[g + 8, 2 * tig])
    acc[off + 3] = acc[off + 3] + float(D[g + 8, 2 * tig + 1])
