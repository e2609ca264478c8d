"""OpenSHMEM-style device API for kernels written in the DSL:  ``from triton_dist.lk import shmem`` and call ``shmem.putmem_signal_block(ctx,
...)`` inside an ``@lk.kernel`` function.

Reference: python/triton_dist/language/extra/cuda/libnvshmem_device.py (the NVSHMEM device API as Triton extern calls: my_pe / n_pes /
team_*, remote_ptr, put / get in thread, warp and block scope with their nbi / rma spellings, put-with-signal, signal_op,
signal_wait_until, fence / quiet, barriers and syncs, broadcast, fcollect).  Here every name is a DSL intrinsic that

* compiles to the C++ function of the same name in ``csrc/td/shmem.cuh`` (namespace ``td::shmem``: plain loads / stores over NVLink to
  ``base + pe * stride`` addresses of the symmetric heap, release / acquire flags, flag-flip team barriers), and
* has a CPU meaning in the interpreter -- the calling thread (or thread 0 of the calling warp / block, with the group's barrier around it)
  performs the same transfer on the emulation backend's shared-memory heap, and the barriers run the same arrival-flag protocol across
  processes -- so a kernel using them is testable under ``kernel.interpret`` with ``world_size > 1`` (tests/dist_worker.py case ``lk``).

Two small structs exist only in device code: ``team = shmem.team_world(ctx)`` / ``shmem.team_split_strided(start, stride, size)`` and
``sync = shmem.make_sync(slots, epoch)`` (``slots``: symmetric ``uint32[2 * world]``, zeroed once; ``epoch``: one local ``uint32``).
Signals are 64-bit words, as in NVSHMEM.  The first argument of everything that touches the heap is the kernel's ``ll.SymmCtx``.
"""
from __future__ import annotations

from types import SimpleNamespace

from . import interp as I
from . import types as T
from .language import Intrinsic, _i
from .types import Pointer, Struct, i32, u32, u64

CMP_EQ, CMP_NE, CMP_GT, CMP_LE, CMP_LT, CMP_GE = range(6)
NVSHMEM_CMP_EQ, NVSHMEM_CMP_NE, NVSHMEM_CMP_GT, NVSHMEM_CMP_LE, NVSHMEM_CMP_LT, NVSHMEM_CMP_GE = range(6)
SIGNAL_SET, SIGNAL_ADD = 9, 10
NVSHMEM_SIGNAL_SET, NVSHMEM_SIGNAL_ADD = 9, 10
NVSHMEM_TEAM_WORLD = 0

Team = Struct("td::shmem::Team", {"start": i32, "stride": i32, "size": i32}, 12)
Sync = Struct("td::shmem::Sync", {"slots": Pointer(u32), "epoch": Pointer(u32)}, 16)


# ---- interpreter helpers --------------------------------------------------------------------------------------------------------
def _group(kind: int):
    """(is the calling thread the one that acts, barrier to run before / after) for thread (0), warp (1) and block (2) scope."""
    c = I.cur()
    if kind == 2:
        return c.linear == 0, I.syncthreads
    if kind == 1:
        return c.linear % 32 == 0, (lambda: I.warp_collect(0) and None)
    return True, (lambda: None)


def _peer(ctx, p, pe):
    import triton_dist.utils as U
    t = U.symm_at(p.owner, int(pe))
    return I.Ptr(t.view(-1), p.off, p.elem, owner=t)


def _copy_bytes(dst, src, nbytes):
    n = int(nbytes)
    d, s = dst.reinterpret(T.u8), src.reinterpret(T.u8)
    d.base[d.off:d.off + n] = s.base[s.off:s.off + n]


def _put(kind):
    def f(ctx, dst, src, nbytes, pe):
        act, bar = _group(kind)
        bar()
        if act:
            _copy_bytes(_peer(ctx, dst, pe), src, nbytes)
        bar()
    return f


def _get(kind):
    def f(ctx, dst, src, nbytes, pe):
        act, bar = _group(kind)
        bar()
        if act:
            _copy_bytes(dst, _peer(ctx, src, pe), nbytes)
        bar()
    return f


def _sig_tensor(p):
    return p.owner.view(-1)[p.off:]


def _interp_signal_op(ctx, sig, value, op, pe):
    from ..language import shmem as H
    H.signal_op(_sig_tensor(sig), int(value), int(op), int(pe))


def _interp_signal_wait_until(sig, cmp, value):
    from ..language import shmem as H
    return H.signal_wait_until(sig.base[sig.off:], int(cmp), int(value))


def _put_signal(kind):
    def f(ctx, dst, src, nbytes, sig, sig_val, sig_op, pe):
        act, bar = _group(kind)
        bar()
        if act:
            _copy_bytes(_peer(ctx, dst, pe), src, nbytes)
            _interp_signal_op(ctx, sig, sig_val, sig_op, pe)
        bar()
    return f


def _team_index_of(t, pe):
    d = int(pe) - t.start
    if d < 0 or d % t.stride or d // t.stride >= t.size:
        return -1
    return d // t.stride


def _sync(kind):
    def f(ctx, team, s):
        from .. import language as dl
        act, bar = _group(kind)
        bar()
        if act:
            epoch = (int(s.epoch[0]) + 1) & 0xFFFFFFFF
            arr = s.slots + (epoch & 1) * ctx.world
            if _team_index_of(team, ctx.rank) >= 0:
                members = [team.start + i * team.stride for i in range(team.size)]
                mine = arr.owner.view(-1)[arr.off + ctx.rank:]
                for pe in members:
                    dl.notify(mine, pe, epoch, "set")
                for pe in members:
                    dl.wait(arr.base[arr.off + pe:], 1, wait_value=epoch, geq=True)
            s.epoch[0] = epoch
        bar()
    return f


def _sync_all(kind):
    g = _sync(kind)
    return lambda ctx, s: g(ctx, SimpleNamespace(start=0, stride=1, size=ctx.world), s)


def _broadcast(kind, typed):
    def f(ctx, team, s, dst, src, n, root):
        act, bar = _group(kind)
        nbytes = int(n) * (dst.elem.nbytes if typed else 1)
        bar()
        if act and _team_index_of(team, ctx.rank) == int(root):
            for i in range(team.size):
                _copy_bytes(_peer(ctx, dst, team.start + i * team.stride), src, nbytes)
        bar()
        _sync(kind)(ctx, team, s)
    return f


def _fcollect(kind):
    def f(ctx, team, s, dst, src, n):
        act, bar = _group(kind)
        nbytes = int(n) * dst.elem.nbytes
        me = _team_index_of(team, ctx.rank)
        bar()
        if act and me >= 0:
            for q in range(team.size):
                i = (me + q) % team.size
                d = _peer(ctx, dst, team.start + i * team.stride).reinterpret(T.u8)
                _copy_bytes(d + me * nbytes, src, nbytes)
        bar()
        _sync(kind)(ctx, team, s)
    return f


def _ns(n):
    return "td::shmem::" + n


def _args(n):
    return ", ".join("{%d}" % i for i in range(n))


def _def(name, ret, nargs, interp=None, doc=""):
    return _i("shmem_" + name, ret, f"{_ns(name)}({_args(nargs)})", nargs, interp=interp, doc=doc)


# ---- PEs and teams --------------------------------------------------------------------------------------------------------------
my_pe = _def("my_pe", i32, 1, lambda ctx: ctx.rank)
n_pes = _def("n_pes", i32, 1, lambda ctx: ctx.world)
team_world = _def("team_world", Team, 1, lambda ctx: SimpleNamespace(start=0, stride=1, size=ctx.world))
team_split_strided = _i("shmem_team_split_strided", Team, "td::shmem::Team{{{0}, {1}, {2}}}", 3,
                        interp=lambda start, stride, size: SimpleNamespace(start=int(start), stride=int(stride), size=int(size)),
                        doc="(start, stride, size) -> the team {start, start + stride, ...} of `size` PEs")
team_n_pes = _def("team_n_pes", i32, 1, lambda t: t.size)
team_my_pe = _def("team_my_pe", i32, 2, lambda ctx, t: _team_index_of(t, ctx.rank))
team_pe = _def("team_pe", i32, 2, lambda t, idx: t.start + int(idx) * t.stride)
team_translate_pe = _def("team_translate_pe", i32, 3, lambda st, pe, dt: _team_index_of(dt, st.start + int(pe) * st.stride)
                         if 0 <= int(pe) < st.size else -1)
make_sync = _i("shmem_make_sync", Sync, "td::shmem::Sync{{{0}, {1}}}", 2, interp=lambda slots, epoch: SimpleNamespace(slots=slots, epoch=epoch),
               doc="(slots: symmetric uint32[2 * world], epoch: local uint32[1]) -> the state the barriers / collectives advance")
remote_ptr = _i("shmem_remote_ptr", lambda a: a[1].ty, "td::shmem::remote_ptr({0}, {1}, {2})", 3, interp=_peer)
remote_mc_ptr = _i("shmem_remote_mc_ptr", lambda a: a[1].ty, "td::shmem::remote_mc_ptr({0}, {1})", 2)

# ---- ordering -------------------------------------------------------------------------------------------------------------------
fence = _def("fence", None, 0, lambda: None)
quiet = _def("quiet", None, 0, lambda: None)


def _interp_int_p(ctx, dst, value, pe):
    _peer(ctx, dst, pe)[0] = int(value)


int_p = _def("int_p", None, 4, _interp_int_p)

# ---- put / get (thread, warp, block scope; nbi and rma spellings are the same transfers here: a store is already non-blocking) --------
putmem, putmem_warp, putmem_block = (_def(n, None, 5, _put(k)) for k, n in enumerate(("putmem", "putmem_warp", "putmem_block")))
getmem, getmem_warp, getmem_block = (_def(n, None, 5, _get(k)) for k, n in enumerate(("getmem", "getmem_warp", "getmem_block")))
putmem_nbi, putmem_nbi_warp, putmem_nbi_block = putmem, putmem_warp, putmem_block
getmem_nbi, getmem_nbi_warp, getmem_nbi_block = getmem, getmem_warp, getmem_block
putmem_rma, putmem_rma_warp, putmem_rma_block = putmem, putmem_warp, putmem_block
putmem_rma_nbi, putmem_rma_nbi_warp, putmem_rma_nbi_block = putmem, putmem_warp, putmem_block

# ---- signals --------------------------------------------------------------------------------------------------------------------
signal_op = _def("signal_op", None, 5, _interp_signal_op, "(ctx, sig, value, SIGNAL_SET | SIGNAL_ADD, pe): release store / add on pe's signal word")
signal_wait_until = _def("signal_wait_until", u64, 3, _interp_signal_wait_until, "(sig, CMP_*, value) -> the value that satisfied the comparison")
putmem_signal, putmem_signal_warp, putmem_signal_block = (
    _def(n, None, 8, _put_signal(k)) for k, n in enumerate(("putmem_signal", "putmem_signal_warp", "putmem_signal_block")))
putmem_signal_nbi, putmem_signal_nbi_warp, putmem_signal_nbi_block = putmem_signal, putmem_signal_warp, putmem_signal_block
putmem_signal_rma, putmem_signal_rma_warp, putmem_signal_rma_block = putmem_signal, putmem_signal_warp, putmem_signal_block
putmem_signal_rma_nbi, putmem_signal_rma_nbi_warp, putmem_signal_rma_nbi_block = putmem_signal, putmem_signal_warp, putmem_signal_block

# ---- barriers / syncs -----------------------------------------------------------------------------------------------------------
team_sync, team_sync_warp, team_sync_block = (_def(n, None, 3, _sync(k)) for k, n in enumerate(("team_sync", "team_sync_warp", "team_sync_block")))
barrier, barrier_warp, barrier_block = (_def(n, None, 3, _sync(k)) for k, n in enumerate(("barrier", "barrier_warp", "barrier_block")))
sync_all, sync_all_warp, sync_all_block = (_def(n, None, 2, _sync_all(k)) for k, n in enumerate(("sync_all", "sync_all_warp", "sync_all_block")))
barrier_all, barrier_all_warp, barrier_all_block = (
    _def(n, None, 2, _sync_all(k)) for k, n in enumerate(("barrier_all", "barrier_all_warp", "barrier_all_block")))

# ---- collectives ----------------------------------------------------------------------------------------------------------------
broadcastmem, broadcastmem_warp, broadcastmem_block = (
    _def(n, None, 7, _broadcast(k, False)) for k, n in enumerate(("broadcastmem", "broadcastmem_warp", "broadcastmem_block")))
broadcast, broadcast_warp, broadcast_block = (
    _def(n, None, 7, _broadcast(k, True)) for k, n in enumerate(("broadcast", "broadcast_warp", "broadcast_block")))
fcollect, fcollect_warp, fcollect_block = (_def(n, None, 6, _fcollect(k)) for k, n in enumerate(("fcollect", "fcollect_warp", "fcollect_block")))

__all__ = [n for n, v in list(globals().items()) if isinstance(v, Intrinsic) or n.isupper() or n in ("Team", "Sync")]
