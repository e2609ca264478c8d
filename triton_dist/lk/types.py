"""Type system of the kernel DSL.

The reference's little_kernel has its own type lattice (python/little_kernel/core/type_system.py) that a type-inference pass walks
before code generation.  Here types are small Python objects that serve three roles at once: annotations in a kernel signature
(``x: ll.ptr[ll.f32]``), cast functions inside kernel bodies (``ll.u32(tid)``), and -- in the CPU interpreter -- value wrappers with
C wrap-around semantics.
"""
from __future__ import annotations

import struct
from typing import Any, Optional, Sequence, Tuple


class Type:
    cname: str = "void"

    def __repr__(self):
        return f"<lk {self.cname}>"


class VoidType(Type):
    cname = "void"


class Scalar(Type):
    """kind: 'i' signed, 'u' unsigned, 'f' float, 'b' bool.  Calling a scalar type casts (compile time: folded; interpreter: wraps)."""

    def __init__(self, name: str, cname: str, bits: int, kind: str, torch_name: Optional[str] = None):
        self.name, self.cname, self.bits, self.kind, self.torch_name = name, cname, bits, kind, torch_name

    @property
    def is_int(self):
        return self.kind in "iu"

    @property
    def is_float(self):
        return self.kind == "f"

    @property
    def is_half(self):
        return self.kind == "f" and self.bits == 16

    @property
    def nbytes(self):
        return max(1, self.bits // 8)

    def wrap(self, v):
        """Python value -> the value a C variable of this type would hold."""
        if self.kind == "b":
            return bool(v)
        if self.kind == "f":
            v = float(v)
            if self.bits == 32:
                return struct.unpack("f", struct.pack("f", v))[0]
            if self.bits == 16:
                import torch
                return float(torch.tensor(v, dtype=getattr(torch, self.torch_name)))
            return v
        v = int(v) & ((1 << self.bits) - 1)
        if self.kind == "i" and v >> (self.bits - 1):
            v -= 1 << self.bits
        return v

    def __call__(self, v=0):      # interpreter / compile-time cast
        return self.wrap(v)

    def literal(self, v) -> str:
        if self.kind == "b":
            return "true" if v else "false"
        if self.kind == "f":
            r = repr(float(v))
            if r in ("inf", "-inf", "nan"):
                r = {"inf": "INFINITY", "-inf": "-INFINITY", "nan": "NAN"}[r]
                return r if self.bits == 64 else f"(({self.cname}){r})"
            if "e" not in r and "." not in r:
                r += ".0"
            if self.bits == 64:
                return r
            if self.bits == 32:
                return r + "f"
            return f"(({self.cname}){r}f)"
        v = int(v)
        suffix = {("u", 32): "u", ("u", 64): "ull", ("i", 64): "ll"}.get((self.kind, self.bits), "")
        if self.bits < 32:
            return f"(({self.cname}){v})"
        if self.kind == "i" and self.bits == 32 and v == -(1 << 31):
            return "(-2147483647 - 1)"
        return f"{v}{suffix}"


void = VoidType()
bool_ = Scalar("bool", "bool", 8, "b", "bool")
i8 = Scalar("i8", "int8_t", 8, "i", "int8")
u8 = Scalar("u8", "uint8_t", 8, "u", "uint8")
i16 = Scalar("i16", "int16_t", 16, "i", "int16")
u16 = Scalar("u16", "uint16_t", 16, "u", "uint16")
i32 = Scalar("i32", "int", 32, "i", "int32")
u32 = Scalar("u32", "uint32_t", 32, "u", "uint32")
i64 = Scalar("i64", "int64_t", 64, "i", "int64")
u64 = Scalar("u64", "uint64_t", 64, "u", "uint64")
f16 = Scalar("f16", "__half", 16, "f", "float16")
bf16 = Scalar("bf16", "__nv_bfloat16", 16, "f", "bfloat16")
f32 = Scalar("f32", "float", 32, "f", "float32")
f64 = Scalar("f64", "double", 64, "f", "float64")
e4m3 = Scalar("e4m3", "__nv_fp8_e4m3", 8, "f", "float8_e4m3fn")

SCALARS = {t.name: t for t in (bool_, i8, u8, i16, u16, i32, u32, i64, u64, f16, bf16, f32, f64, e4m3)}
# aliases in the reference's spelling (little_kernel/core/type_system.py)
ALIASES = {"int8": i8, "uint8": u8, "int16": i16, "uint16": u16, "int32": i32, "uint32": u32, "int64": i64, "uint64": u64,
           "float16": f16, "bfloat16": bf16, "float32": f32, "float64": f64, "bool_": bool_}


class Pointer(Type):
    def __init__(self, elem: Type, const: bool = False):
        self.elem, self.const = elem, const

    @property
    def cname(self):
        return f"{'const ' if self.const else ''}{self.elem.cname}*"

    def __eq__(self, o):
        return isinstance(o, Pointer) and o.elem == self.elem

    def __hash__(self):
        return hash(("ptr", self.elem))


class _PtrFactory:
    """``ll.ptr[ll.f32]`` / ``ll.Tensor[ll.bf16]``."""

    def __getitem__(self, elem) -> Pointer:
        if not isinstance(elem, Type):
            raise TypeError(f"ptr[...] needs a DSL type, got {elem!r}")
        return Pointer(elem)


ptr = _PtrFactory()
Tensor = ptr
void_ptr = Pointer(void)


class Array(Type):
    """A fixed-shape array in shared / dynamic-shared / local (register) space.  Indexing with a tuple flattens row-major."""

    def __init__(self, elem: Type, shape: Sequence[int], space: str):
        self.elem, self.shape, self.space = elem, tuple(int(s) for s in shape), space

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.numel * getattr(self.elem, "nbytes", 16)

    @property
    def cname(self):
        return f"{self.elem.cname}*"


class Struct(Type):
    """A C++ struct the generated code can take apart by field (``dim3``, ``uint4``, ``td::SymmCtx``)."""

    def __init__(self, cname: str, fields: dict, nbytes: int = 16):
        self.cname, self.fields, self.nbytes = cname, dict(fields), nbytes


uint4 = Struct("uint4", {"x": u32, "y": u32, "z": u32, "w": u32})
float4 = Struct("float4", {"x": f32, "y": f32, "z": f32, "w": f32})
float2 = Struct("float2", {"x": f32, "y": f32}, 8)
uint2 = Struct("uint2", {"x": u32, "y": u32}, 8)
SymmCtx = Struct("td::SymmCtx", {"rank": i32, "world": i32, "base": u64, "stride": u64, "mc_base": u64}, 32)


class TmaDescriptorType(Type):
    """``CUtensorMap`` passed ``const __grid_constant__`` (the only way TMA accepts it from kernel parameters)."""
    cname = "CUtensorMap"


TmaDescriptor = TmaDescriptorType()


class _Wrapper:
    def __init__(self, tag):
        self.tag = tag

    def __getitem__(self, t):
        return t          # const[...] / grid_constant[...] are accepted for source compatibility; the code generator decides


const = _Wrapper("const")
grid_constant = _Wrapper("grid_constant")


class ConstExpr:
    """Annotation for compile-time values: ``def f(x, N: ll.constexpr)`` / ``n: ll.constexpr = BM // 2``."""


constexpr = ConstExpr()
template = constexpr


# ------------------------------------------------------------------------------------------------------------
def promote(a: Type, b: Type) -> Type:
    """Result type of ``a (op) b`` following the C usual arithmetic conversions (16-bit floats compute in fp32)."""
    if isinstance(a, Pointer):
        return a
    if isinstance(b, Pointer):
        return b
    if not isinstance(a, Scalar) or not isinstance(b, Scalar):
        raise TypeError(f"cannot combine {a} and {b}")
    if a.is_float or b.is_float:
        if (a.is_float and a.bits == 64) or (b.is_float and b.bits == 64):
            return f64
        return f32
    ra, rb = (a if a.bits >= 32 else i32), (b if b.bits >= 32 else i32)      # integer promotion
    if ra.kind == "b":
        ra = i32
    if rb.kind == "b":
        rb = i32
    if ra.bits != rb.bits:
        return ra if ra.bits > rb.bits else rb
    return ra if ra.kind == "u" else rb


def type_of_const(v: Any) -> Scalar:
    if isinstance(v, bool):
        return bool_
    if isinstance(v, int):
        if -(1 << 31) <= v < (1 << 31):
            return i32
        if 0 <= v < (1 << 32):
            return u32
        if -(1 << 63) <= v < (1 << 63):
            return i64
        return u64
    if isinstance(v, float):
        return f32
    raise TypeError(f"no DSL type for constant {v!r}")


def from_torch_dtype(dt) -> Scalar:
    name = str(dt).replace("torch.", "")
    for t in SCALARS.values():
        if t.torch_name == name:
            return t
    raise TypeError(f"no DSL type for {dt}")


def shape_tuple(s) -> Tuple[int, ...]:
    if isinstance(s, int):
        return (s,)
    return tuple(int(x) for x in s)
