"""The reference's CUDA helper vocabulary (language/extra/cuda/language_extra.py: ``tid``, ``ld``, ``st``, ``atomic_add``, ``atomic_cas``,
``__syncthreads``, ``multimem_*``, ``ld_acquire``, ``red_release``, ``__shfl_*``, ...) as DSL intrinsics, so a kernel that was written
against those names ports to an ``@lk.kernel`` function by changing its import:

    from triton_dist.lk import language_extra as le          # reference: from triton_dist.language.extra import language_extra
    ...
    if le.tid(0) == 0:
        le.st(flag, 1, scope="sys", semantic="release")

Memory-model helpers take ``scope`` (``"cta" | "gpu" | "sys"``) and ``semantic`` (``"relaxed" | "acquire" | "release" | "acq_rel"``)
as compile-time strings, exactly as the reference does, and lower to ONE scoped PTX instruction (``ld.acquire.sys.b32`` ...) whose operand
width follows the pointer's element type.  Everything also has a CPU meaning in the interpreter (one Python thread per CUDA thread; a
lock stands in for atomicity, program order for the fences).
"""
from __future__ import annotations

from . import interp as I
from . import language as ll
from . import types as T
from .language import Intrinsic, _i
from .types import Pointer, Scalar, i32, u32, u64
from .values import CompileError, Val

_SCOPES = {"cta": "cta", "block": "cta", "gpu": "gpu", "device": "gpu", "sys": "sys", "system": "sys"}
_SEMS = ("relaxed", "acquire", "release", "acq_rel", "weak", "volatile")


def _const_str(cg, v, node, what, default):
    if v is None:
        return default
    if not v.is_const or not isinstance(v.const, str):
        raise CompileError(f"language_extra: '{what}' must be a compile-time string", node, cg)
    return v.const


def _scope_sem(cg, args, kwargs, node, n_pos, d_scope, d_sem):
    extra = list(args[n_pos:])
    scope = kwargs.get("scope", extra[0] if extra else None)
    sem = kwargs.get("semantic", extra[1] if len(extra) > 1 else None)
    scope = _SCOPES.get(_const_str(cg, scope, node, "scope", d_scope))
    sem = _const_str(cg, sem, node, "semantic", d_sem)
    if scope is None or sem not in _SEMS:
        raise CompileError("language_extra: scope is cta / gpu / sys, semantic is relaxed / acquire / release / acq_rel", node, cg)
    return scope, sem


def _elem(cg, v, node):
    t = v.ty
    if not isinstance(t, Pointer) or not isinstance(t.elem, Scalar) or t.elem.bits not in (32, 64):
        raise CompileError("language_extra: scoped memory operations take a pointer to a 32- or 64-bit scalar", node, cg)
    return t.elem


def _reg(e: Scalar):
    """(asm constraint letter, typeless PTX suffix) of a 32- / 64-bit scalar."""
    if e.is_float:
        return ("d" if e.bits == 64 else "f"), "b" + str(e.bits)
    return ("l" if e.bits == 64 else "r"), "b" + str(e.bits)


def _helper(cg, name: str, src: str) -> str:
    """Define a ``__device__`` helper once per generated translation unit (calls stay plain expressions: no GNU statement expressions,
    which nvcc's front end does not accept as operands of another asm statement)."""
    key = ("language_extra", name)
    if key not in cg.module.device_keys:
        cg.module.device_keys[key] = (name, T.void)
        cg.module.device_src.append(src)
    return name


def _emit_ld(cg, args, kwargs, node):
    scope, sem = _scope_sem(cg, args, kwargs, node, 1, "gpu", "relaxed")
    e = _elem(cg, args[0], node)
    c, suffix = _reg(e)
    if sem in ("weak", "volatile"):
        q = "ld.volatile.global" if sem == "volatile" else "ld.weak.global"
    else:
        if sem not in ("relaxed", "acquire"):
            raise CompileError("language_extra.ld: semantic is relaxed or acquire", node, cg)
        q = f"ld.{sem}.{scope}.global"
    fn = _helper(cg, f"le_ld_{sem}_{scope}_{e.name}",
                 f'__device__ __forceinline__ {e.cname} le_ld_{sem}_{scope}_{e.name}(const {e.cname}* p) {{ {e.cname} v; '
                 f'asm volatile("{q}.{suffix} %0, [%1];" : "={c}"(v) : "l"(p) : "memory"); return v; }}')
    return Val(f"{fn}({cg.rvalue(args[0])})", e)


def _emit_st(cg, args, kwargs, node):
    scope, sem = _scope_sem(cg, args, kwargs, node, 2, "gpu", "relaxed")
    e = _elem(cg, args[0], node)
    c, suffix = _reg(e)
    if sem not in ("relaxed", "release"):
        raise CompileError("language_extra.st: semantic is relaxed or release", node, cg)
    fn = _helper(cg, f"le_st_{sem}_{scope}_{e.name}",
                 f'__device__ __forceinline__ void le_st_{sem}_{scope}_{e.name}({e.cname}* p, {e.cname} v) {{ '
                 f'asm volatile("st.{sem}.{scope}.global.{suffix} [%0], %1;" ::"l"(p), "{c}"(v) : "memory"); }}')
    return Val(f"{fn}({cg.rvalue(args[0])}, ({e.cname})({cg.rvalue(args[1])}))", T.void)


def _arith_type(e: Scalar) -> str:
    # 64-bit signed adds are issued as .u64 (two's complement: the same bits; PTX has no atom.add.s64)
    return ("f" if e.is_float else ("s" if e.kind == "i" and e.bits == 32 else "u")) + str(e.bits)


def _emit_atomic(op):
    def emit(cg, args, kwargs, node):
        n_pos = 3 if op == "cas" else 2
        scope, sem = _scope_sem(cg, args, kwargs, node, n_pos, "gpu", "relaxed")
        if sem not in ("relaxed", "acquire", "release", "acq_rel"):
            raise CompileError(f"language_extra.atomic_{op}: semantic is relaxed / acquire / release / acq_rel", node, cg)
        e = _elem(cg, args[0], node)
        c, _ = _reg(e)
        p = cg.rvalue(args[0])
        if op == "cas":
            fn = _helper(cg, f"le_cas_{sem}_{scope}_{e.name}",
                         f'__device__ __forceinline__ {e.cname} le_cas_{sem}_{scope}_{e.name}({e.cname}* p, {e.cname} cmp, {e.cname} val) {{ {e.cname} v; '
                         f'asm volatile("atom.{sem}.{scope}.global.cas.b{e.bits} %0, [%1], %2, %3;" : "={c}"(v) : "l"(p), "{c}"(cmp), "{c}"(val) : "memory"); '
                         f'return v; }}')
            return Val(f"{fn}({p}, ({e.cname})({cg.rvalue(args[1])}), ({e.cname})({cg.rvalue(args[2])}))", e)
        fn = _helper(cg, f"le_atom_{op}_{sem}_{scope}_{e.name}",
                     f'__device__ __forceinline__ {e.cname} le_atom_{op}_{sem}_{scope}_{e.name}({e.cname}* p, {e.cname} a) {{ {e.cname} v; '
                     f'asm volatile("atom.{sem}.{scope}.global.{op}.{_arith_type(e)} %0, [%1], %2;" : "={c}"(v) : "l"(p), "{c}"(a) : "memory"); return v; }}')
        return Val(f"{fn}({p}, ({e.cname})({cg.rvalue(args[1])}))", e)
    return emit


def _emit_red(cg, args, kwargs, node):
    scope, sem = _scope_sem(cg, args, kwargs, node, 2, "gpu", "release")
    if sem not in ("relaxed", "release"):
        raise CompileError("language_extra.red_release: semantic is relaxed or release", node, cg)
    e = _elem(cg, args[0], node)
    c, _ = _reg(e)
    fn = _helper(cg, f"le_red_{sem}_{scope}_{e.name}",
                 f'__device__ __forceinline__ void le_red_{sem}_{scope}_{e.name}({e.cname}* p, {e.cname} a) {{ '
                 f'asm volatile("red.{sem}.{scope}.global.add.{_arith_type(e)} [%0], %1;" ::"l"(p), "{c}"(a) : "memory"); }}')
    return Val(f"{fn}({cg.rvalue(args[0])}, ({e.cname})({cg.rvalue(args[1])}))", T.void)


def _emit_fence(cg, args, kwargs, node):
    sem = _const_str(cg, kwargs.get("semantic", args[0] if args else None), node, "semantic", "acq_rel")
    scope = _SCOPES.get(_const_str(cg, kwargs.get("scope", args[1] if len(args) > 1 else None), node, "scope", "gpu"))
    if sem not in ("acq_rel", "sc") or scope is None:
        raise CompileError("language_extra.fence: semantic is acq_rel or sc, scope is cta / gpu / sys", node, cg)
    return Val(f'asm volatile("fence.{sem}.{scope};" ::: "memory")', T.void)


def _emit_membar(cg, args, kwargs, node):
    scope = _SCOPES.get(_const_str(cg, kwargs.get("scope", args[0] if args else None), node, "scope", "gpu"))
    if scope is None:
        raise CompileError("language_extra.membar: scope is cta / gpu / sys", node, cg)
    level = {"cta": "cta", "gpu": "gl", "sys": "sys"}[scope]
    return Val(f'asm volatile("membar.{level};" ::: "memory")', T.void)


def _dim(name):
    def emit(cg, args, kwargs, node):
        a = args[0]
        if not a.is_const or a.const not in (0, 1, 2):
            raise CompileError(f"language_extra.{name}: the axis is a compile-time 0 / 1 / 2", node, cg)
        return Val(f"((int){'threadIdx' if name == 'tid' else 'blockDim'}.{'xyz'[a.const]})", i32)
    return emit


# ---- interpreter meanings -------------------------------------------------------------------------------------------------------
def _i_ld(p, scope="gpu", semantic="relaxed"):
    return p[0]


def _i_st(p, v, scope="gpu", semantic="relaxed"):
    p[0] = v


def _i_rmw(fn):
    return lambda p, v, scope="gpu", semantic="relaxed": I.atomic_rmw(p, 0, lambda o: fn(o, v))


# ---- the vocabulary -------------------------------------------------------------------------------------------------------------
__syncthreads = ll.syncthreads
tid = Intrinsic("le_tid", None, emit=_dim("tid"), interp=lambda axis: I.cur().tid[int(axis)], doc="tid(axis) -> threadIdx.{x,y,z}")
ntid = Intrinsic("le_ntid", None, emit=_dim("ntid"), interp=lambda axis: I.cur().bdim[int(axis)], doc="ntid(axis) -> blockDim.{x,y,z}")
laneid = ll.lane_id
smid = ll.smid
globaltimer = globaltimer_lo = ll.globaltimer
ld = Intrinsic("le_ld", None, emit=_emit_ld, interp=_i_ld, doc="ld(ptr, scope='gpu', semantic='relaxed') -> *ptr with one scoped PTX load")
st = Intrinsic("le_st", None, emit=_emit_st, interp=_i_st, doc="st(ptr, value, scope='gpu', semantic='relaxed')")
ld_acquire = Intrinsic("le_ld_acquire", None, emit=lambda cg, a, k, n: _emit_ld(cg, a[:1], {**k, "semantic": Val("", None, "acquire"),
                                                                              **({"scope": a[1]} if len(a) > 1 else {})}, n),
                       interp=lambda p, scope="gpu": p[0], doc="ld_acquire(ptr, scope='gpu')")
atomic_add = Intrinsic("le_atomic_add", None, emit=_emit_atomic("add"), interp=_i_rmw(lambda o, v: o + v),
                       doc="atomic_add(ptr, value, scope='gpu', semantic='relaxed') -> old value")
atomic_cas = Intrinsic("le_atomic_cas", None, emit=_emit_atomic("cas"),
                       interp=lambda p, c, v, scope="gpu", semantic="relaxed": I.atomic_rmw(p, 0, lambda o: v if o == c else o),
                       doc="atomic_cas(ptr, compare, value, scope='gpu', semantic='relaxed') -> old value")
red_release = Intrinsic("le_red_release", None, emit=_emit_red,
                        interp=lambda p, v, scope="gpu", semantic="release": I.atomic_rmw(p, 0, lambda o: o + v) and None,
                        doc="red_release(ptr, value, scope='gpu'): fire-and-forget release add")
arrive_inc = red_release
fence = Intrinsic("le_fence", None, emit=_emit_fence, interp=lambda semantic="acq_rel", scope="gpu": None, doc="fence(semantic='acq_rel', scope='gpu')")
membar = Intrinsic("le_membar", None, emit=_emit_membar, interp=lambda scope="gpu": None, doc="membar(scope='gpu')")
__fence = membar
tma_sync = ll.fence_proxy_async


def _interp_wait_eq(p, value, scope="sys", semantic="acquire"):
    ll._interp_wait(p, 1, value, False)


def _emit_wait_eq(cg, args, kwargs, node):
    fn = _helper(cg, "le_wait_eq_u32", "__device__ __forceinline__ void le_wait_eq_u32(const uint32_t* p, uint32_t v) "
                                       "{ while (td::ptx::ld_acquire_sys(p) != v) {} }")
    return Val(f"{fn}({cg.rvalue(args[0])}, (uint32_t)({cg.rvalue(args[1])}))", T.void)


wait_eq = Intrinsic("le_wait_eq", None, emit=_emit_wait_eq, interp=_interp_wait_eq,
             doc="wait_eq(ptr, value): the calling thread spins (acquire, system scope) until *ptr == value")


def _interp_add_per_warp(p, v, scope="gpu", semantic="relaxed"):
    lane = I.cur().linear % 32
    old = I.atomic_rmw(p, 0, lambda o: o + v) if lane == 0 else 0
    return I.warp_exchange(old, lambda l: 0)


atomic_add_per_warp = _i("le_atomic_add_per_warp", u32,
                         "__shfl_sync(0xffffffffu, (td::ptx::lane_id() == 0 ? atomicAdd({0}, {1}) : 0u), 0)", 2, interp=_interp_add_per_warp,
                         doc="lane 0 adds, every lane of the warp gets the old value")
__shfl_sync_i32 = _i("le_shfl_sync", i32, "__shfl_sync({0}, {1}, {2})", 3, interp=lambda mask, v, src: I.warp_exchange(v, lambda l: int(src)))
__shfl_up_sync_i32 = _i("le_shfl_up_sync", i32, "__shfl_up_sync({0}, {1}, {2})", 3, interp=lambda mask, v, d: I.warp_exchange(v, lambda l: l - int(d)))
__shfl_down_sync_i32 = _i("le_shfl_down_sync", i32, "__shfl_down_sync({0}, {1}, {2})", 3,
                          interp=lambda mask, v, d: I.warp_exchange(v, lambda l: l + int(d)))
__shfl_xor_sync_i32 = _i("le_shfl_xor_sync", i32, "__shfl_xor_sync({0}, {1}, {2})", 3,
                         interp=lambda mask, v, d: I.warp_exchange(v, lambda l: l ^ int(d)))
__ballot_sync = _i("le_ballot_sync", u32, "__ballot_sync({0}, {1})", 2,
                   interp=lambda mask, pred: sum((1 << i) for i, b in enumerate(I.warp_collect(bool(pred))) if b))
# vector / multicast accesses: the 16-byte forms the DSL already has
ld_vector = load_v4_u32 = ll.ld_v4
st_vector = st_v4_u32 = ll.st_v4
multimem_st_v4 = ll.multimem_st_v4
multimem_ld_reduce_v4 = ll.multimem_ld_reduce_bf16x8
multimem_st_b32 = _i("le_multimem_st_b32", None, 'asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"({0}), "r"((uint32_t)({1})) : "memory")', 2)
multimem_st_b64 = _i("le_multimem_st_b64", None, 'asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"({0}), "l"((uint64_t)({1})) : "memory")', 2)
multimem_st_v2 = _i("le_multimem_st_v2", None,
                    'asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {{%1, %2}};" ::"l"({0}), "r"((uint32_t)({1})), "r"((uint32_t)({2})) : "memory")', 3)
pack_b32_v2 = _i("le_pack_b32_v2", u64, "(((uint64_t)(uint32_t)({1}) << 32) | (uint64_t)(uint32_t)({0}))", 2,
                 interp=lambda lo, hi: ((int(hi) & 0xFFFFFFFF) << 32) | (int(lo) & 0xFFFFFFFF))
pack = pack_b32_v2
unpack_lo = _i("le_unpack_lo", u32, "((uint32_t)({0}))", 1, interp=lambda v: int(v) & 0xFFFFFFFF)
unpack_hi = _i("le_unpack_hi", u32, "((uint32_t)(((uint64_t)({0})) >> 32))", 1, interp=lambda v: (int(v) >> 32) & 0xFFFFFFFF)


def unpack(v):
    """Interpreter-only convenience (the compiled form is ``unpack_lo`` / ``unpack_hi``: DSL calls return one value)."""
    return int(v) & 0xFFFFFFFF, (int(v) >> 32) & 0xFFFFFFFF
