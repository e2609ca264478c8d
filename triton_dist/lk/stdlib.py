"""Small device-side library for DSL kernels: binary searches over sorted arrays, the grid barrier, calls to external C++ functions.

Reference: kernels/nvidia/common_ops.py -- ``bisect_left/right[_aligned]`` device helpers (:264-361: tile-wide searches used by the EP
kernels to map a token index to its expert segment), ``unsafe_barrier_on_this_grid`` / ``cooperative_barrier_on_this_grid`` (:61-130)
-- and language/core.py ``extern_call`` (:85-116).  Here they are plain per-thread device functions / intrinsics of the DSL.
"""
from __future__ import annotations

from . import interp as I
from . import language as ll
from . import types as T
from .compiler import DeviceFunction
from .language import Intrinsic, _i
from .values import CompileError, Val


def _bisect_left(a, n, x):
    """First index i in [0, n] with a[i] >= x (a ascending)."""
    lo = 0
    hi = n
    while lo < hi:
        mid = (lo + hi) // 2
        if a[mid] < x:
            lo = mid + 1
        else:
            hi = mid
    return lo


def _bisect_right(a, n, x):
    """First index i in [0, n] with a[i] > x (a ascending)."""
    lo = 0
    hi = n
    while lo < hi:
        mid = (lo + hi) // 2
        if a[mid] <= x:
            lo = mid + 1
        else:
            hi = mid
    return lo


bisect_left = DeviceFunction(_bisect_left, True)
bisect_right = DeviceFunction(_bisect_right, True)
# the reference's *_aligned forms are the same searches specialised for power-of-two lengths; a scalar loop does not need the distinction
bisect_left_aligned, bisect_right_aligned = bisect_left, bisect_right


def _interp_grid_barrier(counter, target):
    I.syncthreads()
    if I.cur().linear == 0:
        I.atomic_rmw(counter, 0, lambda o: o + 1)
        ll._interp_wait(counter, 1, target, True)
    I.syncthreads()


grid_barrier = _i("grid_barrier", None, "td::grid_barrier({0}, {1})", 2, interp=_interp_grid_barrier,
                  doc="(counter: zero-initialised uint32 in global memory, target = grid size * number of barriers so far): all CTAs of a "
                      "co-resident grid meet (release add + acquire spin by thread 0, __syncthreads around)")
unsafe_barrier_on_this_grid = cooperative_barrier_on_this_grid = grid_barrier


def _emit_extern_call(cg, args, kwargs, node):
    if not args or not args[0].is_const or not isinstance(args[0].const, str):
        raise CompileError("extern_call(symbol, ret_type, *args): the symbol is a compile-time string", node, cg)
    ret = args[1].obj if args[1].obj is not None else args[1].const
    if ret is None:
        ret = T.void
    if not isinstance(ret, T.Type):
        raise CompileError("extern_call: the second argument is the return type (an ll type or None)", node, cg)
    return Val(f"{args[0].const}({', '.join(cg.rvalue(a) for a in args[2:])})", ret)


EXTERN_INTERP: dict = {}        # symbol -> Python callable: what ``extern_call(symbol, ...)`` means in the CPU interpreter


def _interp_extern_call(symbol, ret, *args):
    fn = EXTERN_INTERP.get(symbol)
    if fn is None:
        raise NotImplementedError(f"extern_call({symbol!r}) has no CPU meaning: register one in triton_dist.lk.stdlib.EXTERN_INTERP")
    v = fn(*args)
    return ret.wrap(v) if hasattr(ret, "wrap") and v is not None else v


extern_call = Intrinsic("extern_call", None, emit=_emit_extern_call, interp=_interp_extern_call,
                        doc='extern_call("ns::fn", ll.f32, a, b): call any __device__ function visible to the generated source '
                            "(the headers under csrc/td are included; add more with Kernel(extra_flags=('-include', path)))")
