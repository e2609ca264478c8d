"""``ll`` -- the vocabulary of the kernel DSL: types, thread indices, and the intrinsic table.

Each intrinsic knows how to (a) emit C++ against the device headers of this framework (``csrc/td/ptx.cuh`` -- mbarrier / TMA / tcgen05 /
multimem, ``csrc/td/primitives.cuh`` -- the symmetric-heap primitives) and (b), where it has a CPU meaning, execute in the interpreter.
The reference keeps one Python module of string-template builtins per area (python/little_kernel/language/intrin/{simt,memory,barrier,
tma,umma,cuda_asm}.py); here one table maps names onto functions that already exist (and are hardware-validated) in the C++ headers, so
a DSL kernel and a hand-written kernel share the same PTX wrappers.
"""
from __future__ import annotations

import math
import sys
from typing import Any, Callable, Dict, Optional, Sequence

from . import interp as I
from . import types as T
from .types import (ALIASES, SCALARS, Array, Pointer, Scalar, Struct, SymmCtx, Tensor, TmaDescriptor, bf16, bool_, const,  # noqa: F401
                    constexpr, e4m3, f16, f32, f64, float2, float4, grid_constant, i8, i16, i32, i64, ptr, template, u8, u16, u32, u64,
                    uint2, uint4, void, void_ptr)
from .values import NOCONST, CompileError, Val, const_val

globals().update(ALIASES)          # ll.int32, ll.uint64, ll.bfloat16 ... (the reference's spelling)

INTRINSICS: Dict[str, "Intrinsic"] = {}


class Intrinsic:
    """name -> C++ template + result type (+ optional interpreter implementation).

    ``template`` is a ``str.format`` pattern: ``{0}``, ``{1}`` ... are argument expressions, ``{kw}`` compile-time keyword values (also
    accepted positionally after ``nargs``), ``{args}`` all positional arguments joined by commas."""

    def __init__(self, name: str, ret, template: Optional[str] = None, nargs: Optional[int] = None, kw: Optional[dict] = None,
                 emit: Optional[Callable] = None, interp: Optional[Callable] = None, declares: bool = False, site: bool = False,
                 doc: str = ""):
        self.name, self.ret, self.template, self.nargs, self.kw = name, ret, template, nargs, dict(kw or {})
        self._emit, self._interp, self.declares, self.site = emit, interp, declares, site
        self.__doc__ = doc or (template or name)
        INTRINSICS[name] = self

    def __repr__(self):
        return f"<ll.{self.name}>"

    # -- interpreter -------------------------------------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        if self._interp is None:
            raise NotImplementedError(f"ll.{self.name} has no CPU meaning (GPU only)")
        if self.site:
            kwargs["_site"] = sys._getframe(1).f_lineno
        return self._interp(*args, **kwargs)

    # -- code generation ---------------------------------------------------------------------------------------
    def emit(self, cg, args: Sequence[Val], kwargs: Dict[str, Val], node=None, target: Optional[str] = None) -> Val:
        if self._emit is not None:
            return self._emit(cg, args, kwargs, node, target) if self.declares else self._emit(cg, args, kwargs, node)
        args = list(args)
        kwv = dict(self.kw)
        if self.nargs is not None and len(args) > self.nargs:           # keyword values given positionally
            extra = args[self.nargs:]
            args = args[:self.nargs]
            for k, v in zip([k for k in self.kw], extra):
                kwargs = {**kwargs, k: v}
        for k, v in kwargs.items():
            if k not in kwv:
                raise CompileError(f"ll.{self.name}: unknown keyword '{k}'", node, cg)
            if not v.is_const:
                raise CompileError(f"ll.{self.name}: '{k}' must be a compile-time constant", node, cg)
            kwv[k] = v.const
        if self.nargs is not None and len(args) != self.nargs:
            raise CompileError(f"ll.{self.name} takes {self.nargs} argument(s), got {len(args)}", node, cg)
        codes = [cg.rvalue(a) for a in args]
        fmt = {k: (int(v) if isinstance(v, bool) else v) for k, v in kwv.items()}
        code = self.template.format(*codes, args=", ".join(codes), **fmt)
        ret = self.ret(args) if callable(self.ret) and not isinstance(self.ret, T.Type) else self.ret
        return Val(code, ret if ret is not None else T.void)


def _i(name, ret, template, nargs=None, kw=None, interp=None, **k):
    return Intrinsic(name, ret, template, nargs, kw, interp=interp, **k)


def _promote_all(args):
    t = args[0].ty
    for a in args[1:]:
        t = T.promote(t, a.ty)
    return t


def _same(args):
    return args[0].ty


def _elem(args):
    t = args[0].ty
    return t.elem if isinstance(t, (Pointer, Array)) else t


# ------------------------------------------------------------------------------------------------------------
# thread / block indices
# ------------------------------------------------------------------------------------------------------------
class Dim3Proxy:
    """``ll.threadIdx.x`` -- C++ builtin when compiled (cast to int: Python-like signed arithmetic), TLS lookup when interpreted."""

    def __init__(self, cname: str, field: str):
        self.cname, self._field = cname, field

    def _get(self, k):
        return getattr(I.cur(), self._field)[k]

    x = property(lambda s: s._get(0))
    y = property(lambda s: s._get(1))
    z = property(lambda s: s._get(2))


threadIdx = Dim3Proxy("threadIdx", "tid")
blockIdx = Dim3Proxy("blockIdx", "bid")
blockDim = Dim3Proxy("blockDim", "bdim")
gridDim = Dim3Proxy("gridDim", "gdim")

for _n, _p, _k in (("threadIdx", threadIdx, "tid"), ("blockIdx", blockIdx, "bid"), ("blockDim", blockDim, "bdim"), ("gridDim", gridDim, "gdim")):
    for _j, _c in enumerate("xyz"):            # the reference's function spelling: ll.threadIdx_x()
        _i(f"{_n}_{_c}", i32, f"((int){_n}.{_c})", 0, interp=(lambda k=_k, j=_j: getattr(I.cur(), k)[j]))

lane_id = _i("lane_id", i32, "((int)td::ptx::lane_id())", 0, interp=lambda: I.cur().linear % 32)
warp_id = _i("warp_id", i32, "((int)(threadIdx.x >> 5))", 0, interp=lambda: I.cur().linear // 32)
smid = _i("smid", i32, "((int)td::ptx::smid())", 0, interp=lambda: 0)
globaltimer = _i("globaltimer", u64, "td::ptx::globaltimer()", 0, interp=lambda: __import__("time").perf_counter_ns())
clock64 = _i("clock64", i64, "clock64()", 0, interp=lambda: __import__("time").perf_counter_ns())
syncthreads = _i("syncthreads", None, "__syncthreads()", 0, interp=I.syncthreads)
sync_threads = _i("sync_threads", None, "__syncthreads()", 0, interp=I.syncthreads)
syncwarp = _i("syncwarp", None, "__syncwarp()", 0, interp=lambda: I.warp_collect(0) and None)
elect_one = _i("elect_one", bool_, "td::ptx::elect_one_sync()", 0, interp=lambda: I.cur().linear % 32 == 0,
               doc="true in exactly one lane of the (converged) warp")
named_bar_sync = _i("named_bar_sync", None, "td::ptx::named_bar_sync({0}, {1})", 2)
named_bar_arrive = _i("named_bar_arrive", None, "td::ptx::named_bar_arrive({0}, {1})", 2)
nanosleep = _i("nanosleep", None, "__nanosleep({0})", 1, interp=lambda ns: __import__("time").sleep(ns * 1e-9))
trap = _i("trap", None, "__trap()", 0, interp=lambda: (_ for _ in ()).throw(RuntimeError("ll.trap()")))
setmaxnreg_inc = _i("setmaxnreg_inc", None, 'asm volatile("setmaxnreg.inc.sync.aligned.u32 {0};")', 1, interp=lambda n: None)
setmaxnreg_dec = _i("setmaxnreg_dec", None, 'asm volatile("setmaxnreg.dec.sync.aligned.u32 {0};")', 1, interp=lambda n: None)

shfl_xor = _i("shfl_xor", _same, "__shfl_xor_sync(0xffffffffu, {0}, {1})", 2, interp=lambda v, m: I.warp_exchange(v, lambda l: l ^ int(m)))
shfl_down = _i("shfl_down", _same, "__shfl_down_sync(0xffffffffu, {0}, {1})", 2, interp=lambda v, d: I.warp_exchange(v, lambda l: l + int(d)))
shfl_up = _i("shfl_up", _same, "__shfl_up_sync(0xffffffffu, {0}, {1})", 2,
             interp=lambda v, d: I.warp_exchange(v, lambda l: l - int(d) if l - int(d) >= 0 else -1))
shfl_idx = _i("shfl_idx", _same, "__shfl_sync(0xffffffffu, {0}, {1})", 2, interp=lambda v, s: I.warp_exchange(v, lambda l: int(s)))
shfl_xor_sync, shfl_down_sync, shfl_up_sync = shfl_xor, shfl_down, shfl_up
ballot = _i("ballot", u32, "__ballot_sync(0xffffffffu, {0})", 1,
            interp=lambda p: sum((1 << i) for i, v in enumerate(I.warp_collect(bool(p))) if v))
warp_any = _i("warp_any", bool_, "(__any_sync(0xffffffffu, {0}) != 0)", 1, interp=lambda p: any(I.warp_collect(bool(p))))
warp_all = _i("warp_all", bool_, "(__all_sync(0xffffffffu, {0}) != 0)", 1, interp=lambda p: all(I.warp_collect(bool(p))))
popc = _i("popc", i32, "__popc({0})", 1, interp=lambda v: bin(int(v) & 0xFFFFFFFF).count("1"))
clz = _i("clz", i32, "__clz({0})", 1, interp=lambda v: 32 - (int(v) & 0xFFFFFFFF).bit_length())
ffs = _i("ffs", i32, "__ffs({0})", 1, interp=lambda v: ((int(v) & -int(v)).bit_length()))

# clusters
cluster_rank = _i("cluster_rank", i32, "((int)td::ptx::cluster_ctarank())", 0, interp=lambda: 0)
cluster_size = _i("cluster_size", i32, "((int)td::ptx::cluster_nctarank())", 0, interp=lambda: 1)
cluster_sync = _i("cluster_sync", None, "td::ptx::cluster_sync()", 0, interp=I.syncthreads)
cluster_arrive = _i("cluster_arrive", None, "td::ptx::cluster_arrive()", 0)
cluster_wait = _i("cluster_wait", None, "td::ptx::cluster_wait()", 0)
mapa = _i("mapa", u32, "td::ptx::mapa({0}, {1})", 2, doc="shared::cta address -> shared::cluster address of CTA {1}")


# ------------------------------------------------------------------------------------------------------------
# memory spaces: declarations
# ------------------------------------------------------------------------------------------------------------
def _decl_args(cg, args, kwargs, node, what):
    vals = list(args)
    shape = vals[0] if vals else kwargs.get("shape")
    dtype = vals[1] if len(vals) > 1 else kwargs.get("dtype")
    if shape is None or dtype is None or not shape.is_const or not isinstance(dtype.obj, T.Type):
        raise CompileError(f"ll.{what}(shape, dtype): shape must be compile-time and dtype a DSL type", node, cg)
    align = kwargs.get("align")
    return T.shape_tuple(shape.const), dtype.obj, (int(align.const) if align is not None else None)


def _emit_shared(cg, args, kwargs, node, target):
    shape, dtype, align = _decl_args(cg, args, kwargs, node, "shared")
    arr = Array(dtype, shape, "shared")
    cg.declare_array(target, arr, align or max(16, getattr(dtype, "nbytes", 16)))
    return Val(target, arr, lvalue=True)


def _emit_dyn_shared(cg, args, kwargs, node, target):
    shape, dtype, align = _decl_args(cg, args, kwargs, node, "dyn_shared")
    arr = Array(dtype, shape, "dyn_shared")
    cg.declare_dyn_shared(target, arr, align or max(16, getattr(dtype, "nbytes", 16)))
    return Val(target, arr, lvalue=True)


def _emit_local(cg, args, kwargs, node, target):
    shape, dtype, _ = _decl_args(cg, args, kwargs, node, "local")
    arr = Array(dtype, shape, "local")
    cg.declare_array(target, arr, None)
    return Val(target, arr, lvalue=True)


def _emit_empty(cg, args, kwargs, node, target):
    """The reference's spelling: ``ll.empty(shape, dtype=, scope="dynamic_shared" | "shared" | "local")``."""
    scope = kwargs.pop("scope", None)
    scope = scope.const if scope is not None else "local"
    fn = {"dynamic_shared": _emit_dyn_shared, "shared": _emit_shared, "local": _emit_local}.get(scope)
    if fn is None:
        raise CompileError(f"ll.empty: unknown scope {scope!r}", node, cg)
    return fn(cg, args, kwargs, node, target)


def _interp_dyn_shared(shape, dtype, align=None, _site=None):
    c = I.cur()
    shape = T.shape_tuple(shape)
    a = align or max(16, getattr(dtype, "nbytes", 16))
    nbytes = int(math.prod(shape)) * getattr(dtype, "nbytes", 16)
    with c.block.lock:          # every thread carves the same layout; the storage is the block's
        off = (c.dyn_off + a - 1) // a * a
        c.dyn_off = off + nbytes
    return I.shared_array(("dyn", off, getattr(dtype, "name", "?")), shape, dtype)


shared = Intrinsic("shared", None, emit=_emit_shared, declares=True, site=True,
                   interp=lambda shape, dtype, align=None, _site=None: I.shared_array(("st", _site), shape, dtype))
dyn_shared = Intrinsic("dyn_shared", None, emit=_emit_dyn_shared, declares=True, site=True, interp=_interp_dyn_shared)
local = Intrinsic("local", None, emit=_emit_local, declares=True,
                  interp=lambda shape, dtype, align=None: I.shared_array(None, shape, dtype, per_block=False))
empty = Intrinsic("empty", None, emit=_emit_empty, declares=True, site=True,
                  interp=lambda shape, dtype=None, scope="local", align=None, _site=None:
                  (_interp_dyn_shared(shape, dtype, align, _site) if scope == "dynamic_shared" else
                   I.shared_array(("st", _site), shape, dtype) if scope == "shared" else I.shared_array(None, shape, dtype, per_block=False)))


def _emit_align_memory(cg, args, kwargs, node):
    cg.align_dyn_shared(int(args[0].const))
    return Val("", T.void)


def _interp_align(a, scope="dynamic_shared"):
    c = I.cur()
    c.dyn_off = (c.dyn_off + a - 1) // a * a


align_memory = Intrinsic("align_memory", None, emit=_emit_align_memory, interp=_interp_align)


# ------------------------------------------------------------------------------------------------------------
# casts / addresses
# ------------------------------------------------------------------------------------------------------------
def _emit_val_cast(cg, args, kwargs, node):
    v, ty = args[0], args[1].obj
    return cg.cast(v, ty)


def _emit_ptr_cast(cg, args, kwargs, node):
    v, ty = args[0], args[1].obj
    if isinstance(ty, Scalar):
        ty = Pointer(ty)
    return Val(f"reinterpret_cast<{ty.cname}>({cg.rvalue(v)})", ty)


def _interp_ptr_cast(p, ty):
    return p.reinterpret(ty) if hasattr(p, "reinterpret") else p


val_cast = Intrinsic("val_cast", None, emit=_emit_val_cast, interp=lambda v, ty: ty.wrap(v))
to = val_cast
ptr_cast = Intrinsic("ptr_cast", None, emit=_emit_ptr_cast, interp=_interp_ptr_cast)
smem_addr = _i("smem_addr", u32, "td::ptx::smem_u32({0})", 1, doc="generic pointer -> 32-bit shared-window address")
cvta_generic_to_shared = smem_addr
addr_of = Intrinsic("addr_of", None, emit=lambda cg, a, k, n: Val(f"(&{a[0].code})", Pointer(a[0].ty)))
float_as_uint = _i("float_as_uint", u32, "__float_as_uint({0})", 1,
                   interp=lambda v: __import__("struct").unpack("I", __import__("struct").pack("f", v))[0])
uint_as_float = _i("uint_as_float", f32, "__uint_as_float({0})", 1,
                   interp=lambda v: __import__("struct").unpack("f", __import__("struct").pack("I", int(v) & 0xFFFFFFFF))[0])
pack_bf16x2 = _i("pack_bf16x2", u32, "td::ptx::pack_bf16x2({0}, {1})", 2)
bf16_lo = _i("bf16_lo", f32, "td::ptx::bf16_lo({0})", 1)
bf16_hi = _i("bf16_hi", f32, "td::ptx::bf16_hi({0})", 1)
bf16_to_float = _i("bf16_to_float", f32, "__bfloat162float({0})", 1, interp=float)
float_to_bf16 = _i("float_to_bf16", bf16, "__float2bfloat16_rn({0})", 1, interp=bf16.wrap)


def _emit_sizeof(cg, args, kwargs, node):
    t = args[0].obj if args[0].obj is not None else args[0].ty
    return const_val(getattr(t, "nbytes", 0))


sizeof = Intrinsic("sizeof", None, emit=_emit_sizeof, interp=lambda t: t.nbytes)

# ------------------------------------------------------------------------------------------------------------
# math
# ------------------------------------------------------------------------------------------------------------
for _n, _c, _f in (("exp", "__expf", math.exp), ("exp2", "exp2f", lambda x: 2.0 ** x), ("log", "__logf", math.log), ("log2", "__log2f", math.log2),
                   ("sqrt", "sqrtf", math.sqrt), ("rsqrt", "rsqrtf", lambda x: 1.0 / math.sqrt(x)), ("fabs", "fabsf", abs),
                   ("floor", "floorf", math.floor), ("ceil", "ceilf", math.ceil), ("tanh", "tanhf", math.tanh), ("sin", "__sinf", math.sin),
                   ("cos", "__cosf", math.cos), ("rcp", "__frcp_rn", lambda x: 1.0 / x), ("ex2_approx", "td::ptx::ex2_approx", lambda x: 2.0 ** x)):
    globals()[_n] = _i(_n, f32, _c + "((float)({0}))", 1, interp=_f)
fma = _i("fma", f32, "fmaf({0}, {1}, {2})", 3, interp=lambda a, b, c: a * b + c)
cdiv = _i("cdiv", _promote_all, "((({0}) + ({1}) - 1) / ({1}))", 2, interp=lambda a, b: (a + b - 1) // b)
min_val = _i("min_val", _promote_all, "min({0}, {1})", 2, interp=min)
max_val = _i("max_val", _promote_all, "max({0}, {1})", 2, interp=max)


# ------------------------------------------------------------------------------------------------------------
# global / shared memory accesses with explicit semantics
# ------------------------------------------------------------------------------------------------------------
def _ld(p, i=0):
    return p[i]


def _st(p, v):
    p[0] = v


ld_acquire_sys = _i("ld_acquire_sys", _elem, "td::ptx::ld_acquire_sys({0})", 1, interp=_ld)
ld_acquire_gpu = _i("ld_acquire_gpu", _elem, "td::ptx::ld_acquire_gpu({0})", 1, interp=_ld)
ld_relaxed_sys = _i("ld_relaxed_sys", _elem, "td::ptx::ld_relaxed_sys({0})", 1, interp=_ld)
ld_volatile = _i("ld_volatile", _elem, "td::ptx::ld_volatile({0})", 1, interp=_ld)
st_release_sys = _i("st_release_sys", None, "td::ptx::st_release_sys({0}, {1})", 2, interp=_st)
st_release_gpu = _i("st_release_gpu", None, "td::ptx::st_release_gpu({0}, {1})", 2, interp=_st)
st_relaxed_sys = _i("st_relaxed_sys", None, "td::ptx::st_relaxed_sys({0}, {1})", 2, interp=_st)
red_add_release_sys = _i("red_add_release_sys", None, "td::ptx::red_release_sys_add({0}, {1})", 2,
                         interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: o + v) and None)
red_add_release_gpu = _i("red_add_release_gpu", None, "td::ptx::red_release_gpu_add({0}, {1})", 2,
                         interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: o + v) and None)
atomic_add = _i("atomic_add", _elem, "atomicAdd({0}, {1})", 2, interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: o + v))
atomic_max = _i("atomic_max", _elem, "atomicMax({0}, {1})", 2, interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: max(o, v)))
atomic_min = _i("atomic_min", _elem, "atomicMin({0}, {1})", 2, interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: min(o, v)))
atomic_exch = _i("atomic_exch", _elem, "atomicExch({0}, {1})", 2, interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: v))
atomic_cas = _i("atomic_cas", _elem, "atomicCAS({0}, {1}, {2})", 3, interp=lambda p, c, v: I.atomic_rmw(p, 0, lambda o: v if o == c else o))
atom_add_acq_rel_sys = _i("atom_add_acq_rel_sys", u32, "td::ptx::atom_add_acq_rel_sys({0}, {1})", 2,
                          interp=lambda p, v: I.atomic_rmw(p, 0, lambda o: o + v))
atom_cas_acq_rel_sys = _i("atom_cas_acq_rel_sys", u32, "td::ptx::atom_cas_acq_rel_sys({0}, {1}, {2})", 3,
                          interp=lambda p, c, v: I.atomic_rmw(p, 0, lambda o: v if o == c else o))
fence_sys = _i("fence_sys", None, "td::ptx::fence_acq_rel_sys()", 0, interp=lambda: None)
fence_gpu = _i("fence_gpu", None, "td::ptx::fence_acq_rel_gpu()", 0, interp=lambda: None)
threadfence = _i("threadfence", None, "__threadfence()", 0, interp=lambda: None)
threadfence_system = _i("threadfence_system", None, "__threadfence_system()", 0, interp=lambda: None)
ld_v4 = _i("ld_v4", uint4, "td::ptx::ld_v4({0})", 1, doc="16-byte global load")
ld_nc_v4 = _i("ld_nc_v4", uint4, "td::ptx::ld_nc_v4({0})", 1, doc="16-byte read-only, no-L1-allocate load")
ld_relaxed_sys_v4 = _i("ld_relaxed_sys_v4", uint4, "td::ptx::ld_relaxed_sys_v4({0})", 1)
st_v4 = _i("st_v4", None, "td::ptx::st_v4({0}, {1})", 2)
st_na_v4 = _i("st_na_v4", None, "td::ptx::st_na_v4({0}, {1})", 2)
ld_shared_v4 = _i("ld_shared_v4", uint4, "td::ptx::ld_shared_v4({0})", 1)
st_shared_v4 = _i("st_shared_v4", None, "td::ptx::st_shared_v4({0}, {1})", 2)
make_uint4 = _i("make_uint4", uint4, "make_uint4({0}, {1}, {2}, {3})", 4,
                interp=lambda x, y, z, w: __import__("types").SimpleNamespace(x=u32(x), y=u32(y), z=u32(z), w=u32(w)))
ldg = _i("ldg", _elem, "__ldg({0})", 1, interp=_ld)
prefetch_l2 = _i("prefetch_l2", None, "td::ptx::prefetch_l2_bulk({0}, {1})", 2)
# NVLS multicast
# NVLS in the interpreter: ``symm_mc`` returns a pointer that remembers it is the multicast alias (``McPtr``); a multimem load reads the word
# from EVERY rank's copy on the emulation heap and adds (what the switch does), a multimem store / reduction touches every copy.
class McPtr:
    def __init__(self, p):
        self.p = p

    def __add__(self, n):
        return McPtr(self.p + n)

    __radd__ = __add__

    def __sub__(self, n):
        return McPtr(self.p - n)

    def copies(self):
        import triton_dist.utils as U
        out = []
        for r in range(U.world_size()):
            t = U.symm_at(self.p.owner, r)
            out.append(I.Ptr(t.view(-1), self.p.off, self.p.elem, owner=t))
        return out


def _need_mc(mc, what):
    if not isinstance(mc, McPtr):
        raise TypeError(f"ll.{what} takes the multicast alias of a symmetric pointer (ll.symm_mc(ctx, p)), got an ordinary pointer")
    return mc


def _interp_mm_ld_reduce(kind):
    def f(mc):
        import struct as _st
        import types as _ty
        from . import pipeline as P
        words = [P.ld_v4(c) for c in _need_mc(mc, "multimem_ld_reduce").copies()]
        out = []
        for k in ("x", "y", "z", "w"):
            ws = [getattr(w, k) for w in words]
            if kind == "f32":
                tot = sum(_st.unpack("f", _st.pack("I", w))[0] for w in ws)
                out.append(_st.unpack("I", _st.pack("f", tot))[0])
            elif kind == "bf16":
                out.append(P.pack_bf16x2(sum(P._bf16_val(w) for w in ws), sum(P._bf16_val(w >> 16) for w in ws)))
            else:
                import torch
                h = lambda b: torch.tensor([b & 0xFFFF], dtype=torch.int32).to(torch.int16).view(torch.float16).float().item()
                lo, hi = sum(h(w) for w in ws), sum(h(w >> 16) for w in ws)
                pk = lambda v: int(torch.tensor([v], dtype=torch.float16).view(torch.int16).item()) & 0xFFFF
                out.append(pk(lo) | (pk(hi) << 16))
        return _ty.SimpleNamespace(x=out[0], y=out[1], z=out[2], w=out[3])
    return f


def _interp_mm_st_v4(mc, v):
    from . import pipeline as P
    for c in _need_mc(mc, "multimem_st_v4").copies():
        P.st_v4(c, v)


def _interp_mm_red_add_u32(mc, v):
    for c in _need_mc(mc, "multimem_red_add_u32").copies():
        I.atomic_rmw(c, 0, lambda o: (o + int(v)) & 0xFFFFFFFF)


multimem_ld_reduce_bf16x8 = _i("multimem_ld_reduce_bf16x8", uint4, "td::ptx::multimem_ld_reduce_bf16x8({0})", 1, interp=_interp_mm_ld_reduce("bf16"))
multimem_ld_reduce_f16x8 = _i("multimem_ld_reduce_f16x8", uint4, "td::ptx::multimem_ld_reduce_f16x8({0})", 1, interp=_interp_mm_ld_reduce("f16"))
multimem_ld_reduce_f32x4 = _i("multimem_ld_reduce_f32x4", uint4, "td::ptx::multimem_ld_reduce_f32x4({0})", 1, interp=_interp_mm_ld_reduce("f32"))
multimem_st_v4 = _i("multimem_st_v4", None, "td::ptx::multimem_st_v4({0}, {1})", 2, interp=_interp_mm_st_v4)
multimem_red_add_u32 = _i("multimem_red_add_u32", None, "td::ptx::multimem_red_add_u32({0}, {1})", 2, interp=_interp_mm_red_add_u32)
red_add_bf16x8 = _i("red_add_bf16x8", None, "td::ptx::red_add_bf16x8({0}, {1})", 2)

# ------------------------------------------------------------------------------------------------------------
# mbarrier / TMA
# ------------------------------------------------------------------------------------------------------------
mbar_init = _i("mbar_init", None, "td::ptx::mbar_init({0}, {1})", 2)
mbar_arrive = _i("mbar_arrive", None, "td::ptx::mbar_arrive({0})", 1)
mbar_arrive_cluster = _i("mbar_arrive_cluster", None, "td::ptx::mbar_arrive_cluster({0}, {1})", 2)
mbar_arrive_expect_tx = _i("mbar_arrive_expect_tx", None, "td::ptx::mbar_arrive_expect_tx({0}, {1})", 2)
mbar_expect_tx = _i("mbar_expect_tx", None, "td::ptx::mbar_expect_tx({0}, {1})", 2)
mbar_wait = _i("mbar_wait", None, "td::ptx::mbar_wait({0}, {1})", 2)
mbar_try_wait = _i("mbar_try_wait", bool_, "td::ptx::mbar_try_wait({0}, {1})", 2)
fence_barrier_init = _i("fence_barrier_init", None, "td::ptx::fence_barrier_init()", 0)
fence_proxy_async = _i("fence_proxy_async", None, "td::ptx::fence_proxy_async()", 0)
fence_proxy_async_smem = _i("fence_proxy_async_smem", None, "td::ptx::fence_proxy_async_smem()", 0)
# the reference's names for the same things (little_kernel/language/intrin/barrier.py)
init_smem_barrier, mbarrier_wait, mbarrier_arrive_and_expect_tx, fence_smem_barrier_init = mbar_init, mbar_wait, mbar_arrive_expect_tx, fence_barrier_init

prefetch_tensormap = _i("prefetch_tensormap", None, "td::ptx::prefetch_tensormap(&{0})", 1)
prefetch_tma_descriptor = prefetch_tensormap
tma_load_2d = _i("tma_load_2d", None, "td::ptx::tma_load_2d(&{0}, {1}, {2}, {3}, {4})", 5,
                 doc="(tensormap, mbarrier, smem_dst, c_inner, c_outer)")
tma_load_3d = _i("tma_load_3d", None, "td::ptx::tma_load_3d(&{0}, {1}, {2}, {3}, {4}, {5})", 6)
tma_load_4d = _i("tma_load_4d", None, "td::ptx::tma_load_4d(&{0}, {1}, {2}, {3}, {4}, {5}, {6})", 7)
tma_load_2d_2sm = _i("tma_load_2d_2sm", None, "td::ptx::tma_load_2d_2sm(&{0}, {1}, {2}, {3}, {4})", 5,
                     doc="cta_group::2 load: the mbarrier of the leader CTA of the pair is signalled")
tma_gather4_2d = _i("tma_gather4_2d", None, "td::ptx::tma_gather4_2d(&{0}, {1}, {2}, {3}, {4}, {5}, {6}, {7})", 8)
tma_store_2d = _i("tma_store_2d", None, "td::ptx::tma_store_2d(&{0}, {1}, {2}, {3})", 4)
tma_store_3d = _i("tma_store_3d", None, "td::ptx::tma_store_3d(&{0}, {1}, {2}, {3}, {4})", 5)
tma_reduce_add_2d = _i("tma_reduce_add_2d", None, "td::ptx::tma_reduce_add_2d(&{0}, {1}, {2}, {3})", 4)
bulk_commit = _i("bulk_commit", None, "td::ptx::bulk_commit()", 0)
bulk_wait = _i("bulk_wait", None, "td::ptx::bulk_wait<{n}>()", 0, kw={"n": 0})
bulk_wait_read = _i("bulk_wait_read", None, "td::ptx::bulk_wait_read<{n}>()", 0, kw={"n": 0})
bulk_g2s = _i("bulk_g2s", None, "td::ptx::bulk_g2s({0}, {1}, {2}, {3})", 4, doc="(smem_dst, gmem_src, bytes, mbarrier)")
bulk_s2g = _i("bulk_s2g", None, "td::ptx::bulk_s2g({0}, {1}, {2})", 3)

# ------------------------------------------------------------------------------------------------------------
# tcgen05 / TMEM
# ------------------------------------------------------------------------------------------------------------
tmem_alloc = _i("tmem_alloc", None, "td::ptx::tmem_alloc<{cta_group}>({0}, {1})", 2, kw={"cta_group": 1},
                doc="(smem slot that receives the TMEM address, columns); one full warp executes it")
tmem_relinquish = _i("tmem_relinquish", None, "td::ptx::tmem_relinquish<{cta_group}>()", 0, kw={"cta_group": 1})
tmem_dealloc = _i("tmem_dealloc", None, "td::ptx::tmem_dealloc<{cta_group}>({0}, {1})", 2, kw={"cta_group": 1})
tc_fence_before = _i("tc_fence_before", None, "td::ptx::tc_fence_before()", 0)
tc_fence_after = _i("tc_fence_after", None, "td::ptx::tc_fence_after()", 0)
tcgen05_fence_before, tcgen05_fence_after = tc_fence_before, tc_fence_after
mma_f16 = _i("mma_f16", None, "td::ptx::mma_f16<{cta_group}>({0}, {1}, {2}, {3}, {4})", 5, kw={"cta_group": 1},
             doc="tcgen05.mma kind::f16 (tmem_d, a_desc, b_desc, idesc, accumulate); issue from ONE thread")
mma_f8f6f4 = _i("mma_f8f6f4", None, "td::ptx::mma_f8f6f4<{cta_group}>({0}, {1}, {2}, {3}, {4})", 5, kw={"cta_group": 1})
mma_i8 = _i("mma_i8", None, "td::ptx::mma_i8<{cta_group}>({0}, {1}, {2}, {3}, {4})", 5, kw={"cta_group": 1})
mma_commit = _i("mma_commit", None, "td::ptx::mma_commit({0})", 1, doc="tcgen05.commit -> mbarrier arrive when prior MMAs retire")
mma_commit_2sm = _i("mma_commit_2sm", None, "td::ptx::mma_commit_2sm({0}, {1})", 2)
tmem_ld_32x32b_x32 = _i("tmem_ld_32x32b_x32", None, "td::ptx::tmem_ld_32x32b_x32({0}, *reinterpret_cast<uint32_t (*)[32]>({1}))", 2,
                        doc="(tmem address, uint32[32] local array): lane l of the warp reads row (lane-group base + l), 32 columns")
tmem_ld_32x32b_x16 = _i("tmem_ld_32x32b_x16", None, "td::ptx::tmem_ld_32x32b_x16({0}, *reinterpret_cast<uint32_t (*)[16]>({1}))", 2)
tmem_ld_wait = _i("tmem_ld_wait", None, "td::ptx::tmem_ld_wait()", 0)
make_smem_desc_k128 = _i("make_smem_desc_k128", u64, "td::ptx::make_smem_desc_k128({0})", 1,
                         doc="UMMA shared-memory descriptor: K-major tile, 128-byte swizzle (what a SWIZZLE_128B TMA box produces)")
make_smem_desc_mn128 = _i("make_smem_desc_mn128", u64, "td::ptx::make_smem_desc_mn128({0}, {1})", 2)


mma_m16n8k16_bf16 = _i("mma_m16n8k16_bf16", None, "td::mma_sync::m16n8k16_bf16({0} + ({1}), {2}, {3}, {4}, {5}, {6}, {7})", 8,
                       doc="warp-level mma.sync m16n8k16 bf16 -> fp32: (acc f32 array, offset, a0, a1, a2, a3, b0, b1) with packed bf16x2 registers; "
                           "accumulates the lane's 4 results into acc[offset .. offset + 3]")


def make_idesc(a_fmt: int, b_fmt: int, M: int, N: int, a_mn_major: int = 0, b_mn_major: int = 0) -> int:
    """kind::f16 / f8f6f4 instruction descriptor (fp32 accumulate) -- plain Python, folded at compile time (mirrors
    ``td::ptx::make_idesc``, csrc/td/ptx.cuh)."""
    return (1 << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


# ------------------------------------------------------------------------------------------------------------
# distributed primitives on the symmetric heap (csrc/td/primitives.cuh; reference: distributed_ops.py:53-107)
# ------------------------------------------------------------------------------------------------------------
def _interp_symm_at(ctx, p, peer):
    import triton_dist.utils as U
    t = U.symm_at(p.owner, int(peer))
    return I.Ptr(t.view(-1), p.off, p.elem, owner=t)


def _interp_notify(ctx, flag, peer, value, op="set"):
    from .. import language as dl
    dl.notify(flag.owner.view(-1)[flag.off:], int(peer), int(value), op)


def _interp_wait(flags, n, value, geq=False):
    from .. import language as dl
    dl.wait(flags.base[flags.off:], int(n), wait_value=int(value), geq=bool(geq))
    return value


def _emit_notify(cg, args, kwargs, node):
    op = kwargs.get("op")
    op = op.const if op is not None else (args[4].const if len(args) > 4 else "set")
    c = [cg.rvalue(a) for a in args[:4]]
    return Val(f"td::notify({c[0]}, {c[1]}, {c[2]}, {c[3]}, td::SignalOp::{'ADD' if op == 'add' else 'SET'})", T.void)


def _emit_wait(cg, args, kwargs, node):
    geq = kwargs.get("geq")
    geq = bool(geq.const) if geq is not None else (bool(args[3].const) if len(args) > 3 else False)
    c = [cg.rvalue(a) for a in args[:3]]
    return Val(f"td::wait<{'true' if geq else 'false'}, true>({c[0]}, {c[1]}, {c[2]})", u32)


rank = _i("rank", i32, "td::rank({0})", 1, interp=lambda ctx: ctx.rank)
num_ranks = _i("num_ranks", i32, "td::num_ranks({0})", 1, interp=lambda ctx: ctx.world)
symm_at = _i("symm_at", lambda a: a[1].ty if isinstance(a[1].ty, Pointer) else Pointer(a[1].ty.elem), "td::symm_at({0}, {1}, {2})", 3,
             interp=_interp_symm_at, doc="(ctx, local pointer, peer) -> the same offset in peer's heap segment")
symm_mc = _i("symm_mc", lambda a: a[1].ty, "td::symm_mc({0}, {1})", 2, interp=lambda ctx, p: McPtr(p),
             doc="(ctx, local pointer) -> NVLS multicast alias (loads reduce over all ranks' copies in the switch, stores reach all of them)")
notify = Intrinsic("notify", None, emit=_emit_notify, interp=_interp_notify,
                   doc="(ctx, flag, peer, value, op='set'|'add'): release store / add of a flag on peer; call from ONE thread")
wait = Intrinsic("wait", None, emit=_emit_wait, interp=_interp_wait,
                 doc="(flags, n, value, geq=False): the calling WARP spins (acquire, system scope) until n flags match")
wait_ge = _i("wait_ge", None, "td::wait_ge<true>({0}, {1})", 2, interp=lambda f, v: _interp_wait(f, 1, v, True) and None)
consume_token = _i("consume_token", _same, "td::consume_token({0}, {1})", 2, interp=lambda v, t: v)
barrier_all_block = _i("barrier_all_block", None, "td::barrier_all_block({0}, {1}, {2})", 3)
putmem_block = _i("putmem_block", None, "td::putmem_block({0}, {1}, {2}, {3}, {4})", 5)
getmem_block = _i("getmem_block", None, "td::getmem_block({0}, {1}, {2}, {3}, {4})", 5)
putmem_warp = _i("putmem_warp", None, "td::putmem_warp({0}, {1}, {2}, {3}, {4})", 5)
getmem_warp = _i("getmem_warp", None, "td::getmem_warp({0}, {1}, {2}, {3}, {4})", 5)
putmem_signal_block = _i("putmem_signal_block", None, "td::putmem_signal_block({0}, {1}, {2}, {3}, {4}, {5}, {6})", 7)


# ------------------------------------------------------------------------------------------------------------
# escape hatches
# ------------------------------------------------------------------------------------------------------------
def _emit_asm(cg, args, kwargs, node):
    """``ll.asm("red.release.sys.global.add.u32 [%0], %1;", inputs=[p, v])`` -- constraint letters come from the operand types."""
    text = args[0].const
    outs = kwargs.get("outputs")
    ins = kwargs.get("inputs")
    outs = outs.const if outs is not None else []
    ins = ins.const if ins is not None else []
    mem = kwargs.get("memory")
    mem = True if mem is None else bool(mem.const)

    def letter(v: Val):
        t = v.ty
        if isinstance(t, (Pointer, Array)):
            return "l"
        if isinstance(t, Scalar):
            if t.is_float:
                return "d" if t.bits == 64 else "f"
            return {8: "r", 16: "h", 32: "r", 64: "l"}[t.bits]
        raise CompileError("ll.asm: unsupported operand type", node, cg)
    o = ", ".join(f'"={letter(v)}"({v.code})' for v in outs)
    i = ", ".join(f'"{letter(v)}"({cg.rvalue(v)})' for v in ins)
    esc = text.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n")
    clob = ' : "memory"' if mem else ""
    return Val(f'asm volatile("{esc}" : {o} : {i}{clob})', T.void)


asm = Intrinsic("asm", None, emit=_emit_asm)


def _emit_printf(cg, args, kwargs, node):
    fmt = args[0].const.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n")
    rest = "".join(", " + cg.rvalue(a) for a in args[1:])
    return Val(f'printf("{fmt}"{rest})', T.void)


printf = Intrinsic("printf", None, emit=_emit_printf, interp=lambda fmt, *a: print(fmt % a, end=""))


def unroll(r, factor: Optional[int] = None):
    """``for i in ll.unroll(range(n))`` -> ``#pragma unroll``."""
    return r


def static_range(*a):
    """``for i in ll.static_range(n)``: unrolled by the code generator; ``i`` is a compile-time constant in the body."""
    return range(*a)


# ------------------------------------------------------------------------------------------------------------
# CPU meaning of the asynchronous-pipeline intrinsics: the functional model of ``pipeline.py`` (mbarrier phases and transaction
# counts, TMA tile loads, tensor memory, cta_group::1 / ::2 MMAs and commits) -- the GEMM ladder runs in the interpreter
# ------------------------------------------------------------------------------------------------------------
def _install_pipeline_model():
    from . import pipeline as P
    noop = lambda *a, **k: None      # noqa: E731
    table = {
        "mbar_init": P.mbar_init, "mbar_arrive": P.mbar_arrive, "mbar_arrive_cluster": P.mbar_arrive_cluster,
        "mbar_arrive_expect_tx": P.mbar_arrive_expect_tx, "mbar_expect_tx": P.mbar_expect_tx, "mbar_wait": P.mbar_wait,
        "mbar_try_wait": P.mbar_try_wait, "fence_barrier_init": noop, "fence_proxy_async": noop, "fence_proxy_async_smem": noop,
        "prefetch_tensormap": noop, "tma_load_2d": P.tma_load_2d, "tma_load_2d_2sm": P.tma_load_2d_2sm,
        "tmem_alloc": P.tmem_alloc, "tmem_relinquish": noop, "tmem_dealloc": noop, "tc_fence_before": noop, "tc_fence_after": noop,
        "mma_f16": P.mma_f16, "mma_commit": P.mma_commit, "mma_commit_2sm": P.mma_commit_2sm,
        "tmem_ld_32x32b_x32": P.tmem_ld_32x32b_x32, "tmem_ld_32x32b_x16": lambda t, r: P.tmem_ld_32x32b_x32(t, r, 16), "tmem_ld_wait": noop,
        "make_smem_desc_k128": P.make_smem_desc_k128, "smem_addr": P.smem_addr, "pack_bf16x2": P.pack_bf16x2, "st_v4": P.st_v4,
        "mma_m16n8k16_bf16": P.mma_m16n8k16_bf16, "ld_v4": P.ld_v4, "ld_nc_v4": P.ld_v4, "st_na_v4": P.st_v4, "red_add_bf16x8": P.red_add_bf16x8,
        "bf16_lo": lambda w: P._bf16_val(w), "bf16_hi": lambda w: P._bf16_val(int(w) >> 16),
        "cluster_sync": P.cluster_sync, "cluster_rank": lambda: I.cur().block.cta_rank, "cluster_size": lambda: len(I.cur().block.cluster),
    }
    for name, fn in table.items():
        INTRINSICS[name]._interp = fn


_install_pipeline_model()
