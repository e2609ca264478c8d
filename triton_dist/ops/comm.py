"""Collectives over the symmetric heap: barrier, fast AllReduce, fast AllGather, memory ops.

GPU: csrc/comm_kernels.cu.  No-GPU: the same protocols executed on the shared-memory heap (emulation).

Reference API being matched:
  * ``create_allreduce_ctx`` / ``all_reduce`` / ``get_auto_allreduce_method``
    (/root/reference/python/triton_dist/kernels/nvidia/allreduce.py:109, :1130-1209, :1102)
  * ``AllReduceMethod`` (/root/reference/python/triton_dist/kernels/allreduce.py:31-49)
  * ``create_fast_allgather_context`` / ``fast_allgather`` (kernels/nvidia/low_latency_allgather.py:968)
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import Optional

import torch

from .. import _C
from .. import utils as U

c_ll, c_ull, c_void_p = C.c_longlong, C.c_ulonglong, C.c_void_p


class SymmArgs(C.Structure):
    _fields_ = [("rank", c_ll), ("world", c_ll), ("base", c_ull), ("stride", c_ull), ("mc_base", c_ull)]


def symm_args() -> SymmArgs:
    r, w, base, stride, mc = U.symm_ctx_fields()
    return SymmArgs(r, w, base, stride, mc)


class _ARArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("method", c_ll), ("dtype", c_ll), ("grid", c_ll), ("in_symm", c_ll),
                ("inp", c_void_p), ("out", c_void_p), ("stage", c_void_p), ("stage2", c_void_p),
                ("stage_bytes", c_ll), ("nbytes", c_ll), ("slots", c_void_p), ("phase", c_void_p)]


class _AGArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("mode", c_ll), ("grid", c_ll), ("inp", c_void_p), ("out", c_void_p),
                ("buf", c_void_p), ("buf_bytes", c_ll), ("shard_bytes", c_ll), ("slots", c_void_p), ("phase", c_void_p)]


_C.register("td_barrier_all", C.c_int, [C.POINTER(SymmArgs), c_void_p, c_void_p, c_void_p])
_C.register("td_allreduce", C.c_int, [C.POINTER(_ARArgs), c_void_p])
_C.register("td_allgather", C.c_int, [C.POINTER(_AGArgs), c_void_p])
_C.register("td_copy", C.c_int, [c_void_p, c_void_p, c_ll, C.c_int, c_void_p])
_C.register("td_fill32", C.c_int, [c_void_p, C.c_uint, c_ll, C.c_int, c_void_p])
_C.register("td_reduce_slabs", C.c_int, [c_void_p, c_void_p, c_ll, C.c_int, C.c_int, C.c_int, c_void_p])

_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _stream(stream=None):
    return c_void_p((stream or torch.cuda.current_stream()).cuda_stream)


# ------------------------------------------------------------------------------------------------------------
# barrier
# ------------------------------------------------------------------------------------------------------------
def barrier_all(ctx: "U.BarrierAllContext", stream=None):
    sa = symm_args()
    _C.check(_C.cuda_lib().td_barrier_all(C.byref(sa), c_void_p(ctx.slots.data_ptr()), c_void_p(ctx.epoch.data_ptr()),
                                          _stream(stream)), "td_barrier_all")


# ------------------------------------------------------------------------------------------------------------
# AllReduce
# ------------------------------------------------------------------------------------------------------------
class AllReduceMethod(enum.Enum):
    """Same names as the reference (kernels/allreduce.py:31-42); aliases map onto our four kernels."""
    Unknown = 0
    OneShot = 1
    TwoShot = 2
    DoubleTree = 3
    OneShot_TMA = 4
    OneShot_Multimem = 5
    TwoShot_Multimem = 6
    TwoShot_Multimem_ST = 7
    AllReduce_Max = 8
    OneShot_LL = 9          # flag-in-data low-latency protocol for tiny messages (csrc/allreduce_ll.cu); opt-in, see docs/status.md


class OverlappingAllReduceMethod(enum.Enum):
    Auto = 0
    Consumer_Load = 1
    Consumer_Multimem = 2
    Consumer_Ring_Reduce = 3


_KERNEL_METHOD = {
    AllReduceMethod.OneShot: 0, AllReduceMethod.OneShot_TMA: 0, AllReduceMethod.DoubleTree: 1,
    AllReduceMethod.TwoShot: 1, AllReduceMethod.OneShot_Multimem: 2, AllReduceMethod.TwoShot_Multimem: 3,
    AllReduceMethod.TwoShot_Multimem_ST: 3,
}


def to_allreduce_method(name) -> AllReduceMethod:
    if isinstance(name, AllReduceMethod):
        return name
    table = {m.name.lower(): m for m in AllReduceMethod}
    key = str(name).lower().replace("-", "_")
    if key not in table:
        raise ValueError(f"unknown allreduce method {name}; choose from {sorted(table)}")
    return table[key]


def get_allreduce_methods():
    return [m for m in AllReduceMethod if m not in (AllReduceMethod.Unknown, AllReduceMethod.AllReduce_Max)]


def get_auto_allreduce_method(nbytes: int) -> AllReduceMethod:
    """Latency-bound messages: one-shot; bandwidth-bound: two-shot.  NVLS variants when the multicast mapping
    exists.  Crossover re-derived for NVLink 5 (reference: 64 KB on H800, allreduce.py:1102-1120): a one-shot
    moves (W-1) x nbytes into every GPU, a two-shot 2 x (W-1)/W x nbytes plus one more barrier (~3 us)."""
    mm = U.is_nvshmem_multimem_supported()
    if nbytes <= 256 * 1024:
        return AllReduceMethod.OneShot_Multimem if mm else AllReduceMethod.OneShot
    return AllReduceMethod.TwoShot_Multimem if mm else AllReduceMethod.TwoShot


get_auto_all_reduce_method = get_auto_allreduce_method


def workspace_bytes_per_in_byte(world_size: int, method) -> int:
    """Staging bytes per input byte.  The reference's push kernels need ``world`` (one-shot: every peer's copy lands in my workspace) or
    2 (two-shot) (allreduce.py:52-60); the kernels here PULL -- every rank stages its own input once and peers read it over NVLink or
    through the multicast mapping -- so the answer is 1 for every method (the parity double-buffer is part of the context, not of the
    per-call budget)."""
    to_allreduce_method(method) if not isinstance(method, AllReduceMethod) else method
    return 1


def get_max_chunk_nbytes(workspace_nbytes: int, world_size: int, method) -> int:
    """Largest message one launch handles; ``all_reduce`` splits longer inputs into chunks of this size."""
    return workspace_nbytes // workspace_bytes_per_in_byte(world_size, method)


def get_tree_parent_and_children(N: int, rank: int):
    """Two complementary binary trees over ``N`` (power of two) ranks, the topology of a double-tree all-reduce (reference:
    allreduce.py ``get_tree_parent_and_children``; NCCL's double binary tree): every rank is an interior node in at most one of them,
    so both trees together use every link in both directions.

    Tree A is the in-order perfect binary tree over labels 1..N-1 (the children of a node with lowest set bit b are ``x -/+ b/2``)
    hung under rank 0; tree B is tree A with every label decreased by one (mod N), i.e. hung under rank N-1.
    Returns ``(parent_a, left_a, right_a, parent_b, left_b, right_b)``, -1 where there is none."""
    assert N >= 2 and N & (N - 1) == 0 and 0 <= rank < N

    def in_tree_a(x):
        if x == 0:                                   # super-root: one child, the root of the perfect tree
            return -1, N // 2, -1
        b = x & -x
        k = (x // b - 1) // 2
        parent = 0 if b == N // 2 else (x + b if k % 2 == 0 else x - b)
        return (parent, x - b // 2, x + b // 2) if b > 1 else (parent, -1, -1)

    pa, la, ra = in_tree_a(rank)
    pb, lb, rb = in_tree_a((rank + 1) % N)
    back = lambda x: -1 if x < 0 else (x - 1) % N
    return pa, la, ra, back(pb), back(lb), back(rb)


@dataclass
class AllReduceContext:
    workspace_nbytes: int
    rank: int
    world_size: int
    local_world_size: int
    stage: torch.Tensor = None      # symmetric uint8 [2 * workspace]
    stage2: torch.Tensor = None
    slots: torch.Tensor = None      # symmetric int32 [grid_max * 2 * world]
    phase: torch.Tensor = None      # local int32 [4]
    grid_max: int = 64
    host_calls: int = 0             # emulation backend

    def finalize(self):
        """Collective: every rank must call it.  Frees symmetric memory only after every rank has drained its stream and
        arrived (a peer may still be writing flags into these slots otherwise, and a recycled offset would see them)."""
        if self.stage is None:
            return
        if self.stage.is_cuda:
            torch.cuda.synchronize()
        U.barrier_all_host()
        for impl in self.__dict__.pop("_dsl", {}).values():            # DSL twins created by TD_ALLREDUCE_DSL=1
            impl.finalize()
        for t in (self.stage, self.stage2, self.slots, getattr(self, "ll_buf", None)):
            if t is not None:
                U.get_heap().free_tensor(t)
        self.stage = self.stage2 = self.slots = None
        self.ll_buf = None

    def symm_input(self, nbytes: int, dtype: torch.dtype) -> torch.Tensor:
        """Zero-copy entry: the staging half the NEXT collective on this context will reduce; a producer may write straight
        into it and pass it as ``x`` (the kernel then skips its staging copy).  The half is chosen from the host mirror of
        the call counter (no device sync); it is only valid until the next collective on this context, and not inside a
        CUDA-graph capture (a replay alternates halves) -- :func:`all_reduce` rejects a stale or offset view."""
        if self.stage.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("symm_input() cannot be used while capturing a CUDA graph: write the device-selected half via "
                               "gemm(out_parity=(ctx.phase, ctx.workspace_nbytes)) and pass the stage base instead")
        par = (self.host_calls + 1) & 1
        es = torch.empty(0, dtype=dtype).element_size()
        return self.stage[par * self.workspace_nbytes: par * self.workspace_nbytes + nbytes].view(dtype)[: nbytes // es]

    def zero_copy_ok(self, ptr: int, device_parity: bool) -> bool:
        """Is ``ptr`` a legal zero-copy input?  Only (a) the stage base when the producer wrote the DEVICE-selected half
        (``gemm(out_parity=...)`` contract) or (b) exactly the half the next call reduces.  Any other pointer into the
        staging area (stale half, offset view) would be silently replaced by the other half's content: raise."""
        base = self.stage.data_ptr()
        if not (base <= ptr < base + 2 * self.workspace_nbytes):
            return False
        if device_parity and ptr == base:
            return True
        par = (self.host_calls + 1) & 1
        if ptr == base + par * self.workspace_nbytes and not (self.stage.is_cuda and torch.cuda.is_current_stream_capturing()):
            return True
        raise ValueError("all_reduce/reduce_scatter: input lies inside the context's staging area but is not the half the next "
                         "call reduces (stale symm_input() view, or a view at an offset); pass a fresh symm_input() or a plain tensor")


def create_allreduce_ctx(workspace_nbytes: int, rank: int, world_size: int, local_world_size: int,
                         grid_max: int = 64) -> AllReduceContext:
    heap = U.get_heap()
    ws = (int(workspace_nbytes) + 1023) // 1024 * 1024
    ctx = AllReduceContext(ws, rank, world_size, local_world_size, grid_max=grid_max)
    ctx.stage = heap.tensor((2 * ws,), torch.uint8)
    ctx.stage2 = heap.tensor((2 * ws,), torch.uint8)
    ctx.slots = heap.tensor((grid_max * 2 * world_size,), torch.int32)
    ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


def _ar_grid(nbytes: int, max_sm: int, grid_max: int) -> int:
    # ~32 KB per CTA keeps every CTA's barrier traffic amortised; small messages use few CTAs (latency)
    g = max(1, min(grid_max, (nbytes + 32767) // 32768))
    if max_sm and max_sm > 0:
        g = min(g, max_sm)
    return g


def all_reduce(x: torch.Tensor, method=AllReduceMethod.Unknown, ctx: AllReduceContext = None,
               output: Optional[torch.Tensor] = None, max_sm: int = -1, straggler_option=None,
               stream=None, device_parity_input: bool = False) -> torch.Tensor:
    """SUM all-reduce of ``x`` over the symmetric team.  Messages larger than the workspace are chunked.
    ``device_parity_input``: ``x`` is the stage base and the producer already wrote the device-selected half."""
    assert ctx is not None, "create_allreduce_ctx() first"
    assert x.is_contiguous()
    if output is None:
        output = torch.empty_like(x)
    nbytes = x.numel() * x.element_size()
    if nbytes % 16:
        raise ValueError("all_reduce: message size must be a multiple of 16 bytes")
    if U.get_bool_env("TD_ALLREDUCE_DSL", False) and x.dtype in (torch.bfloat16, torch.float32) and not device_parity_input \
            and (not x.is_cuda or U.is_nvshmem_multimem_supported()):
        # opt-in A/B switch: the same NVLS algorithms as kernels written in the Python DSL (triton_dist/lk/kernels/allreduce_nvls.py);
        # on the emulation backend they run in the DSL interpreter with the multicast model
        return _all_reduce_dsl(x, method, ctx, output)
    if not x.is_cuda:
        return _all_reduce_host(x, ctx, output)
    if isinstance(method, str):
        method = to_allreduce_method(method)
    if method in (AllReduceMethod.Unknown, None):
        method = get_auto_allreduce_method(nbytes)
    if method in (AllReduceMethod.OneShot_Multimem, AllReduceMethod.TwoShot_Multimem,
                  AllReduceMethod.TwoShot_Multimem_ST) and not U.is_nvshmem_multimem_supported():
        method = AllReduceMethod.OneShot if method == AllReduceMethod.OneShot_Multimem else AllReduceMethod.TwoShot
    if straggler_option and straggler_option[0] == ctx.rank:
        torch.cuda._sleep(int(straggler_option[1]))
    if method == AllReduceMethod.OneShot_LL:
        if nbytes <= _LL_MAX_BYTES:
            return _all_reduce_ll(x, ctx, output, stream)
        method = get_auto_allreduce_method(nbytes)
    lib = _C.cuda_lib()
    xb = x.view(torch.uint8).view(-1)
    ob = output.view(torch.uint8).view(-1)
    heap = U.get_heap()
    off = 0
    while off < nbytes:
        n = min(ctx.workspace_nbytes, nbytes - off)
        a = _ARArgs()
        a.symm = symm_args()
        a.method, a.dtype = _KERNEL_METHOD[method], _DT[x.dtype]
        a.grid = _ar_grid(n, max_sm, ctx.grid_max)
        src = xb[off:off + n]
        a.in_symm = 1 if ctx.zero_copy_ok(src.data_ptr(), device_parity_input) else 0
        a.inp, a.out = src.data_ptr(), ob[off:off + n].data_ptr()
        a.stage, a.stage2, a.stage_bytes, a.nbytes = ctx.stage.data_ptr(), ctx.stage2.data_ptr(), ctx.workspace_nbytes, n
        a.slots, a.phase = ctx.slots.data_ptr(), ctx.phase.data_ptr()
        _C.check(lib.td_allreduce(C.byref(a), _stream(stream)), "td_allreduce")
        ctx.host_calls += 1          # host mirror of the device call counter (every collective launch advances it by one)
        off += n
    return output


_LL_MAX_BYTES = 64 << 10


def _all_reduce_dsl(x: torch.Tensor, method, ctx: "AllReduceContext", output: torch.Tensor) -> torch.Tensor:
    from ..lk.kernels.allreduce_nvls import LkNvlsAllReduce
    if isinstance(method, str):
        method = to_allreduce_method(method)
    nbytes = x.numel() * x.element_size()
    if method in (AllReduceMethod.Unknown, None):
        method = get_auto_allreduce_method(nbytes)
    which = "two_shot" if method in (AllReduceMethod.TwoShot, AllReduceMethod.TwoShot_Multimem, AllReduceMethod.TwoShot_Multimem_ST,
                                     AllReduceMethod.DoubleTree) else "one_shot"
    cache = ctx.__dict__.setdefault("_dsl", {})
    if which not in cache:
        cache[which] = LkNvlsAllReduce(ctx.workspace_nbytes, which, grid=1 if not x.is_cuda else 8)
    impl = cache[which]
    off, xb, ob = 0, x.view(-1), output.view(-1)
    per = impl.max_bytes // x.element_size()
    while off < xb.numel():
        n = min(per, xb.numel() - off)
        impl(xb[off:off + n], out=ob[off:off + n])
        off += n
    return output


class _ARLLArgs(C.Structure):
    _fields_ = [("rank", c_ll), ("world", c_ll), ("base", C.c_ulonglong), ("stride", C.c_ulonglong), ("mc_base", C.c_ulonglong),
                ("inp", c_void_p), ("out", c_void_p), ("buf", c_void_p), ("max_words", c_ll), ("nbytes", c_ll), ("phase", c_void_p),
                ("dtype", c_ll), ("grid", c_ll)]


_C.register("td_allreduce_ll", C.c_int, [C.POINTER(_ARLLArgs), c_void_p])


def _all_reduce_ll(x: torch.Tensor, ctx: AllReduceContext, output: torch.Tensor, stream=None) -> torch.Tensor:
    """Flag-in-data all-reduce (<= 64 KB): one kernel, no fence, no barrier.  Its buffer is created on first use (a collective
    allocation: every rank reaches the first LL call together) and is never touched by the other methods."""
    W = ctx.world_size
    max_words = _LL_MAX_BYTES // 4
    if getattr(ctx, "ll_buf", None) is None:
        ctx.ll_buf = U.get_heap().tensor((2, W, max_words, 2), torch.int32)
        ctx.ll_phase = torch.zeros(4, dtype=torch.int32, device=x.device)
        U.barrier_all_host()
    a = _ARLLArgs()
    r, w, base, stride, mc = U.symm_ctx_fields()
    a.rank, a.world, a.base, a.stride, a.mc_base = r, w, base, stride, mc
    a.inp, a.out, a.buf = x.data_ptr(), output.data_ptr(), ctx.ll_buf.data_ptr()
    a.max_words, a.nbytes, a.phase = max_words, x.numel() * x.element_size(), ctx.ll_phase.data_ptr()
    a.dtype, a.grid = _DT[x.dtype], 0
    _C.check(_C.cuda_lib().td_allreduce_ll(C.byref(a), _stream(stream)), "td_allreduce_ll")
    return output


def reduce_scatter(x: torch.Tensor, ctx: AllReduceContext, output: Optional[torch.Tensor] = None, max_sm: int = -1,
                   stream=None, device_parity_input: bool = False) -> torch.Tensor:
    """SUM reduce-scatter along dim 0: every rank contributes ``x`` ([W * n, ...]) and receives its ``n`` rows.
    Staging (zero-copy if ``x`` is ``ctx.symm_input``) -> per-CTA flag barrier -> each rank pulls and reduces only
    its own slice (NVLS ``multimem.ld_reduce`` when available, P2P loads otherwise)."""
    W = ctx.world_size
    assert x.is_contiguous() and x.shape[0] % W == 0
    rows = x.shape[0] // W
    if output is None:
        output = torch.empty((rows,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    nbytes = x.numel() * x.element_size()
    if nbytes > ctx.workspace_nbytes:
        raise ValueError("reduce_scatter: message larger than the context workspace")
    if not x.is_cuda:
        full = _all_reduce_host(x, ctx, torch.empty_like(x))
        output.copy_(full[ctx.rank * rows:(ctx.rank + 1) * rows])
        return output
    if (nbytes // W) % 16:
        raise ValueError("reduce_scatter: per-rank slice must be a multiple of 16 bytes")
    heap = U.get_heap()
    xb = x.view(torch.uint8).view(-1)
    in_stage = ctx.zero_copy_ok(xb.data_ptr(), device_parity_input)
    def mk(method, grid, inp):
        a = _ARArgs()
        a.symm = symm_args()
        a.method, a.dtype, a.grid, a.in_symm = method, _DT[x.dtype], grid, 1
        a.inp, a.out = inp, output.data_ptr()
        a.stage, a.stage2, a.stage_bytes, a.nbytes = ctx.stage.data_ptr(), ctx.stage2.data_ptr(), ctx.workspace_nbytes, nbytes
        a.slots, a.phase = ctx.slots.data_ptr(), ctx.phase.data_ptr()
        return a
    if not in_stage:
        # stage with a SEPARATE launch (method 6): the reduce kernel's barrier relies on "staging finished before this
        # kernel started"; the staging kernel picks the half from the same device-side parity (graph-replay safe)
        st = mk(6, _ar_grid(nbytes, max_sm, ctx.grid_max), xb.data_ptr())
        _C.check(_C.cuda_lib().td_allreduce(C.byref(st), _stream(stream)), "td_allreduce(stage)")
    a = _ARArgs()
    a.symm = symm_args()
    a.method = 5 if U.is_nvshmem_multimem_supported() else 4
    a.dtype = _DT[x.dtype]
    a.grid = _ar_grid(nbytes // W, max_sm, ctx.grid_max)
    a.in_symm = 1
    a.inp, a.out = ctx.stage.data_ptr(), output.data_ptr()
    a.stage, a.stage2, a.stage_bytes, a.nbytes = ctx.stage.data_ptr(), ctx.stage2.data_ptr(), ctx.workspace_nbytes, nbytes
    a.slots, a.phase = ctx.slots.data_ptr(), ctx.phase.data_ptr()
    _C.check(_C.cuda_lib().td_allreduce(C.byref(a), _stream(stream)), "td_allreduce(reduce_scatter)")
    ctx.host_calls += 1
    return output


def _all_reduce_host(x, ctx: AllReduceContext, output):
    """Emulation: stage -> flag-flip barrier -> sum over peer views (one-shot protocol, chunked like the GPU path)."""
    heap = U.get_heap()
    lib = _C.host_lib()
    xb, ob = x.view(torch.uint8).view(-1), output.view(torch.uint8).view(-1)
    total = xb.numel()
    off = 0
    while off < total:
        n = min(ctx.workspace_nbytes, total - off)
        ctx.host_calls += 1
        par = ctx.host_calls & 1
        st = ctx.stage[par * ctx.workspace_nbytes: par * ctx.workspace_nbytes + n]
        st.copy_(xb[off:off + n])
        rc = lib.tdh_barrier_all(heap._handle, heap.offset_of(ctx.slots), 2 * ctx.host_calls, 60_000_000)
        if rc:
            raise TimeoutError(lib.tdh_last_error().decode())
        acc = torch.zeros(n // x.element_size(), dtype=torch.float32)
        for r in range(heap.world):
            acc += heap.peer_view(st, (heap.rank + r) % heap.world).view(x.dtype).float()
        ob[off:off + n].copy_(acc.to(x.dtype).view(torch.uint8))
        off += n
    return output


# ------------------------------------------------------------------------------------------------------------
# AllGather (small / medium messages)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class FastAllGatherContext:
    max_shard_bytes: int
    rank: int
    world_size: int
    buf: torch.Tensor = None        # symmetric uint8 [2][world * 2 * shard]  (x2 for the LL atoms)
    slots: torch.Tensor = None
    phase: torch.Tensor = None
    grid_max: int = 32
    host_calls: int = 0

    def finalize(self):
        for t in (self.buf, self.slots):
            if t is not None:
                U.get_heap().free_tensor(t)
        self.buf = self.slots = None


def create_fast_allgather_context(max_shard_bytes: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                                  grid_max: int = 32) -> FastAllGatherContext:
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    sb = (int(max_shard_bytes) + 1023) // 1024 * 1024
    ctx = FastAllGatherContext(sb, rank, world_size, grid_max=grid_max)
    ctx.buf = heap.tensor((2 * world_size * 2 * sb,), torch.uint8)
    ctx.slots = heap.tensor((grid_max * 2 * world_size,), torch.int32)
    ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


# kernel modes: 0 pull, 1 push, 2 push-LL (flag-in-data), 3 multimem push (NVLS), 4 multimem LL (NVLS).  The reference's staged
# variants (push_3d, push_numa_2d*: low_latency_allgather.py:74-530) stage through one GPU per NUMA node / node because PCIe or
# inter-node links are the bottleneck there; inside one NVSwitch domain every peer is one hop away, so they map onto the direct modes.
_AG_MODES = {"pull": 0, "push": 1, "push_2d": 1, "push_3d": 1, "push_numa_2d": 1, "push_2d_ll": 2, "ll": 2,
             "push_multimem": 3, "push_2d_multimem": 3, "push_2d_ll_multimem": 4, "ll_multimem": 4, "push_numa_2d_ll": 2}


def fast_allgather(shard: torch.Tensor, ctx: FastAllGatherContext, mode: str = "push", output: Optional[torch.Tensor] = None,
                   stream=None) -> torch.Tensor:
    """Gather ``shard`` from every rank into ``[world, *shard.shape]``.  On a single NVSwitch domain the
    reference's 2d/3d/numa ring variants (low_latency_allgather.py:74-400) all collapse to a direct push."""
    assert shard.is_contiguous()
    W = ctx.world_size
    nbytes = shard.numel() * shard.element_size()
    if nbytes > ctx.max_shard_bytes:
        raise ValueError("shard larger than the context was created for")
    if output is None:
        output = torch.empty((W,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device)
    if not shard.is_cuda:
        return _allgather_host(shard, ctx, output)
    m = _AG_MODES[mode]
    if m in (3, 4) and not (U.is_nvshmem_multimem_supported() and W > 1):
        m = 1 if m == 3 else 2          # no NVLS mapping (or a single rank): the unicast twin of the same protocol
    if m in (0, 1, 3) and nbytes % 16:
        m = 4 if m == 3 else 2
    a = _AGArgs()
    a.symm = symm_args()
    a.mode = m
    a.grid = max(1, min(ctx.grid_max, (nbytes + 16383) // 16384))
    a.inp, a.out = shard.data_ptr(), output.data_ptr()
    a.buf, a.buf_bytes, a.shard_bytes = ctx.buf.data_ptr(), W * 2 * ctx.max_shard_bytes, nbytes
    a.slots, a.phase = ctx.slots.data_ptr(), ctx.phase.data_ptr()
    _C.check(_C.cuda_lib().td_allgather(C.byref(a), _stream(stream)), "td_allgather")
    return output


def _allgather_host(shard, ctx: FastAllGatherContext, output):
    heap = U.get_heap()
    ctx.host_calls += 1
    par = ctx.host_calls & 1
    nbytes = shard.numel() * shard.element_size()
    base = par * (ctx.world_size * 2 * ctx.max_shard_bytes)
    src = shard.view(torch.uint8).view(-1)
    for r in range(heap.world):
        p = (heap.rank + r) % heap.world
        heap.peer_view(ctx.buf, p)[base + heap.rank * nbytes: base + (heap.rank + 1) * nbytes].copy_(src)
    lib = _C.host_lib()
    rc = lib.tdh_barrier_all(heap._handle, heap.offset_of(ctx.slots), ctx.host_calls, 60_000_000)
    if rc:
        raise TimeoutError(lib.tdh_last_error().decode())
    output.view(torch.uint8).view(-1).copy_(ctx.buf[base: base + heap.world * nbytes])
    return output


# ------------------------------------------------------------------------------------------------------------
# memory ops (memory_ops.py)
# ------------------------------------------------------------------------------------------------------------
def copy_tensor(dst: torch.Tensor, src: torch.Tensor, num_sms: int = 0, stream=None):
    """Vectorised device copy (memory_ops.py copy_tensor).  ``num_sms`` > 0 bounds the grid (copy next to a GEMM); 0 = whole GPU."""
    nbytes = src.numel() * src.element_size()
    if not src.is_cuda or nbytes % 16 or not (dst.is_contiguous() and src.is_contiguous()):
        dst.copy_(src)
        return dst
    _C.check(_C.cuda_lib().td_copy(c_void_p(dst.data_ptr()), c_void_p(src.data_ptr()), nbytes, num_sms, _stream(stream)), "td_copy")
    return dst


def fill_tensor(dst: torch.Tensor, value: int, num_sms: int = 32, stream=None):
    if not dst.is_cuda or dst.element_size() != 4 or not dst.is_contiguous():
        dst.fill_(value)
        return dst
    _C.check(_C.cuda_lib().td_fill32(c_void_p(dst.data_ptr()), value & 0xFFFFFFFF, dst.numel(), num_sms, _stream(stream)), "td_fill32")
    return dst


def reduce_tensor(out: torch.Tensor, slabs: torch.Tensor, num_sms: int = 64, stream=None):
    """``out = slabs.sum(0)`` with fp32 accumulation; ``slabs``: [nsrc, *out.shape] contiguous."""
    nbytes = out.numel() * out.element_size()
    if not out.is_cuda or nbytes % 16 or out.dtype not in _DT:
        out.copy_(slabs.float().sum(0).to(out.dtype))
        return out
    _C.check(_C.cuda_lib().td_reduce_slabs(c_void_p(out.data_ptr()), c_void_p(slabs.data_ptr()), nbytes, slabs.shape[0],
                                           _DT[out.dtype], num_sms, _stream(stream)), "td_reduce_slabs")
    return out
