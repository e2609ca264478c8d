"""Python AST -> CUDA C++ for DSL kernels.

The reference's little_kernel runs a pass pipeline over its own IR (constant folding, inlining, type inference, special-struct
materialisation: python/little_kernel/core/passes/*.py) and then a visitor-based code generator (codegen/visitors/*.py).  This
implementation is one typed walk: every expression lowers to a ``Val`` (C++ text + DSL type + compile-time value when known), so
constant folding, dead-branch pruning and static unrolling happen while lowering, Python globals / closures are the compile-time
environment, and helper functions are specialised per argument-type tuple like C++ templates.  nvcc does the rest (the generated code is
ordinary C++ against ``csrc/td/*.cuh``).

Python scoping is kept: every local is declared once at the top of the generated function with the type of its first assignment (or its
annotation), so a name assigned in both arms of an ``if`` is one variable, as in Python.
"""
from __future__ import annotations

import ast
import builtins
import inspect
import textwrap
import types as pytypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

from . import language as ll
from . import types as T
from .values import NOCONST, CompileError, Val, const_val

_RESERVED = {"auto", "register", "int", "float", "double", "char", "long", "short", "unsigned", "signed", "void", "const", "volatile",
             "struct", "union", "enum", "class", "template", "typename", "namespace", "new", "delete", "this", "switch", "case", "default",
             "do", "goto", "static", "extern", "inline", "operator", "private", "public", "protected", "virtual", "bool", "true", "false",
             "asm", "min", "max", "threadIdx", "blockIdx", "blockDim", "gridDim", "warpSize", "lk_dyn", "lk_dyn_raw"}

DIM3 = T.Struct("dim3", {"x": T.i32, "y": T.i32, "z": T.i32})

_BINOP = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/", ast.FloorDiv: "//", ast.Mod: "%", ast.LShift: "<<", ast.RShift: ">>",
          ast.BitAnd: "&", ast.BitOr: "|", ast.BitXor: "^", ast.Pow: "**"}
_CMP = {ast.Eq: "==", ast.NotEq: "!=", ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">="}


def _pyfold(op: str, a, b):
    return {"+": lambda: a + b, "-": lambda: a - b, "*": lambda: a * b, "/": lambda: a / b, "//": lambda: a // b, "%": lambda: a % b,
            "<<": lambda: a << b, ">>": lambda: a >> b, "&": lambda: a & b, "|": lambda: a | b, "^": lambda: a ^ b, "**": lambda: a ** b,
            "==": lambda: a == b, "!=": lambda: a != b, "<": lambda: a < b, "<=": lambda: a <= b, ">": lambda: a > b, ">=": lambda: a >= b}[op]()


class DeviceFunction:
    """A Python function marked ``@lk.device`` (plain functions called from kernels with run-time arguments are wrapped on the fly)."""

    def __init__(self, fn, inline: bool = True):
        self.fn, self.inline = fn, inline
        self.__name__ = fn.__name__
        self.__doc__ = fn.__doc__

    def __call__(self, *a, **k):          # interpreter: it is just the Python function
        return self.fn(*a, **k)


class Module:
    """Everything that ends up in one .cu file: device-function specialisations (deduplicated) + kernels."""

    def __init__(self):
        self.device_src: List[str] = []
        self.device_keys: Dict[Any, Tuple[str, T.Type]] = {}
        self.in_progress: set = set()

    def unique(self, base: str) -> str:
        n = sum(1 for k in self.device_keys.values() if k[0].startswith(base + "__"))
        return f"{base}__{n}"


class FnCompiler:
    def __init__(self, module: Module, pyfn, is_kernel: bool, arg_vals: Optional[Sequence[Val]] = None, cname: Optional[str] = None):
        self.module, self.pyfn, self.is_kernel = module, pyfn, is_kernel
        self.name = pyfn.__name__
        self.cname = cname or self.name
        src = textwrap.dedent(inspect.getsource(pyfn))
        tree = ast.parse(src)
        self.fdef = next(n for n in tree.body if isinstance(n, (ast.FunctionDef,)))
        self.first_line = pyfn.__code__.co_firstlineno
        dec = len(self.fdef.decorator_list)
        self.first_line -= 0 if not dec else 0
        self.pyenv = dict(pyfn.__globals__)
        try:
            self.pyenv.update(inspect.getclosurevars(pyfn).nonlocals)
        except Exception:      # noqa: BLE001
            pass
        self.vars: Dict[str, T.Type] = {}          # run-time locals (incl. parameters)
        self.cvars: Dict[str, str] = {}            # python name -> C identifier
        self.cenv: Dict[str, Any] = {}             # compile-time locals
        self.decls: List[str] = []
        self.out: List[str] = []
        self.ind = 1
        self.params: List[Tuple[str, T.Type]] = []
        self.ret_type: Optional[T.Type] = None
        self.loops: List[str] = []
        self.dyn_off = 0
        self.dyn_align = 16
        self.uses_dyn = False
        self.tmp = 0
        self._bind_params(arg_vals)

    # ------------------------------------------------------------------------------------------------------
    def err(self, msg, node=None):
        return CompileError(msg, node, self)

    def cident(self, name: str) -> str:
        c = self.cvars.get(name)
        if c is None:
            c = name + "_" if name in _RESERVED else name
            self.cvars[name] = c
        return c

    def emit(self, line: str):
        self.out.append("  " * self.ind + line)

    def fresh(self, base="t") -> str:
        self.tmp += 1
        return f"lk_{base}{self.tmp}"

    def _eval_annotation(self, node):
        if node is None:
            return None
        try:
            return eval(compile(ast.Expression(node), "<lk>", "eval"), self.pyenv, dict(self.cenv))       # noqa: S307
        except Exception as e:      # noqa: BLE001
            raise self.err(f"cannot evaluate annotation '{ast.unparse(node)}': {e}", node)

    def _bind_params(self, arg_vals):
        a = self.fdef.args
        if a.vararg or a.kwarg or a.kwonlyargs:
            raise self.err("*args / **kwargs / keyword-only parameters are not supported", self.fdef)
        names = [p.arg for p in a.args]
        defaults = [None] * (len(names) - len(a.defaults)) + list(a.defaults)
        for idx, (p, d) in enumerate(zip(a.args, defaults)):
            ann = getattr(self.pyfn, "__annotations__", {}).get(p.arg)
            if ann is None or isinstance(ann, str):       # strings: ``from __future__ import annotations`` -- evaluate the AST node
                ann = self._eval_annotation(p.annotation)
            v = arg_vals[idx] if arg_vals is not None and idx < len(arg_vals) else None
            if v is None and d is not None and not self.is_kernel:
                v = self.expr(d)
            if isinstance(ann, T.ConstExpr) or (v is not None and v.ty is None):
                if v is None or not (v.is_const or v.obj is not None):
                    raise self.err(f"parameter '{p.arg}' must be a compile-time constant", p)
                self.cenv[p.arg] = v.const if v.is_const else v.obj
                continue
            if isinstance(ann, T.Type):
                ty = ann
            elif v is not None:
                ty = v.ty
            else:
                raise self.err(f"kernel parameter '{p.arg}' needs a type annotation", p)
            if isinstance(ty, T.Array) and self.is_kernel:
                raise self.err("arrays cannot be kernel parameters", p)
            self.vars[p.arg] = ty
            self.params.append((self.cident(p.arg), ty))
        if self.fdef.returns is not None:
            r = self._eval_annotation(self.fdef.returns)
            if isinstance(r, T.Type):
                self.ret_type = r

    # ------------------------------------------------------------------------------------------------------
    # values
    # ------------------------------------------------------------------------------------------------------
    def wrap(self, obj, node=None) -> Val:
        if isinstance(obj, (bool, int, float)):
            return const_val(obj)
        if obj is None or isinstance(obj, (str, list, tuple, dict, range, bytes)):
            return Val("", None, obj)
        if isinstance(obj, ll.Dim3Proxy):
            return Val(obj.cname, DIM3)
        return Val("", None, NOCONST, obj)

    def rvalue(self, v: Val) -> str:
        if v.ty is None:
            if v.is_const and isinstance(v.const, str):
                return '"' + v.const.replace('"', '\\"') + '"'
            raise self.err(f"a compile-time-only value ({v.const if v.is_const else v.obj!r}) is used as a run-time operand")
        return v.code

    def cast(self, v: Val, ty: T.Type) -> Val:
        if v.ty is None:
            raise self.err("cannot cast a compile-time-only object")
        if isinstance(ty, T.Scalar) and v.is_const and isinstance(v.const, (bool, int, float)):
            c = ty.wrap(v.const)
            return Val(ty.literal(c), ty, c)
        if v.ty == ty or (isinstance(ty, T.Scalar) and isinstance(v.ty, T.Scalar) and v.ty.name == ty.name):
            return Val(v.code, ty)
        if isinstance(ty, T.Scalar):
            src = v.ty
            code = v.code
            if isinstance(src, T.Scalar) and src.is_half:
                code, src = f"((float)({code}))", T.f32
            if ty.name == "bf16":
                return Val(f"__float2bfloat16_rn((float)({code}))", ty)
            if ty.name == "f16":
                return Val(f"__float2half_rn((float)({code}))", ty)
            if ty.kind == "b":
                return Val(f"(({code}) != 0)", ty)
            return Val(f"(({ty.cname})({code}))", ty)
        if isinstance(ty, T.Pointer):
            if isinstance(v.ty, (T.Pointer, T.Array)):
                return Val(f"reinterpret_cast<{ty.cname}>({v.code})", ty)
            return Val(f"reinterpret_cast<{ty.cname}>((unsigned long long)({v.code}))", ty)
        return Val(v.code, ty)

    def as_float(self, v: Val) -> str:
        if isinstance(v.ty, T.Scalar) and v.ty.is_half:
            return f"((float)({v.code}))"
        return v.code

    # ------------------------------------------------------------------------------------------------------
    # declarations
    # ------------------------------------------------------------------------------------------------------
    def declare_scalar(self, name: str, ty: T.Type) -> str:
        c = self.cident(name)
        self.vars[name] = ty
        cn = ty.cname
        self.decls.append(f"{cn} {c};")
        return c

    def declare_array(self, name: str, arr: T.Array, align: Optional[int]):
        if name in self.vars:
            raise self.err(f"array '{name}' is declared twice")
        c = self.cident(name)
        self.vars[name] = arr
        if arr.space == "shared":
            self.decls.append(f"__shared__ __align__({align}) {arr.elem.cname} {c}[{arr.numel}];")
        else:
            self.decls.append(f"{arr.elem.cname} {c}[{arr.numel}];")

    def _table_type(self, values) -> T.Scalar:
        if any(isinstance(x, float) for x in values):
            return T.f32
        ty = T.i32
        for x in values:
            t = T.type_of_const(int(x))
            ty = t if t.bits > ty.bits or (t.bits == ty.bits and t.kind == "u") else ty
        return ty

    def const_table(self, values) -> str:
        key = ("tbl", tuple(values))
        name = self.tables.get(key) if hasattr(self, "tables") else None
        if name is None:
            if not hasattr(self, "tables"):
                self.tables = {}
            ty = self._table_type(values)
            name = self.tables[key] = f"lk_tbl{len(self.tables)}"
            self.decls.append(f"const {ty.cname} {name}[{len(values)}] = {{{', '.join(ty.literal(ty.wrap(x)) for x in values)}}};")
        return name

    def align_dyn_shared(self, a: int):
        self.dyn_off = (self.dyn_off + a - 1) // a * a
        self.dyn_align = max(self.dyn_align, a)

    def declare_dyn_shared(self, name: str, arr: T.Array, align: int):
        if not self.is_kernel:
            raise self.err("dynamic shared memory is carved in the kernel body (pass the arrays to helpers)")
        if name in self.vars:
            raise self.err(f"array '{name}' is declared twice")
        self.align_dyn_shared(align)
        c = self.cident(name)
        self.vars[name] = arr
        self.uses_dyn = True
        self.decls.append(f"{arr.elem.cname}* const {c} = reinterpret_cast<{arr.elem.cname}*>(lk_dyn + {self.dyn_off});")
        self.dyn_off += arr.nbytes

    @property
    def dyn_smem_bytes(self) -> int:
        if not self.uses_dyn:
            return 0
        return self.dyn_off + (self.dyn_align if self.dyn_align > 16 else 0)

    # ------------------------------------------------------------------------------------------------------
    # expressions
    # ------------------------------------------------------------------------------------------------------
    def expr(self, node) -> Val:
        m = getattr(self, "e_" + type(node).__name__, None)
        if m is None:
            raise self.err(f"unsupported expression: {type(node).__name__}", node)
        return m(node)

    def e_Constant(self, node):
        return self.wrap(node.value, node) if not isinstance(node.value, (bool, int, float)) else const_val(node.value)

    def e_Name(self, node):
        n = node.id
        if n in self.vars:
            return Val(self.cident(n), self.vars[n], lvalue=True)
        if n in self.cenv:
            return self.wrap(self.cenv[n], node)
        if n in self.pyenv:
            return self.wrap(self.pyenv[n], node)
        if hasattr(builtins, n):
            return self.wrap(getattr(builtins, n), node)
        raise self.err(f"name '{n}' is not defined", node)

    def e_Attribute(self, node):
        b = self.expr(node.value)
        if b.ty is None:
            o = b.const if b.is_const else b.obj
            try:
                return self.wrap(getattr(o, node.attr), node)
            except AttributeError:
                raise self.err(f"{o!r} has no attribute '{node.attr}'", node)
        if b.ty is DIM3:
            return Val(f"((int){b.code}.{node.attr})", T.i32)
        if isinstance(b.ty, T.Struct):
            if node.attr not in b.ty.fields:
                raise self.err(f"{b.ty.cname} has no field '{node.attr}'", node)
            return Val(f"{b.code}.{node.attr}", b.ty.fields[node.attr], lvalue=b.lvalue)
        raise self.err(f"attribute access on a value of type {b.ty}", node)

    def e_Tuple(self, node):
        vals = [self.expr(e) for e in node.elts]
        if all(v.is_const or v.is_obj for v in vals):
            return Val("", None, tuple(v.const if v.is_const else v.obj for v in vals))
        return Val("", None, tuple(vals))            # run-time operands (ll.asm(inputs=[...]))

    def e_List(self, node):
        v = self.e_Tuple(node)
        return Val("", None, list(v.const))

    def e_UnaryOp(self, node):
        v = self.expr(node.operand)
        if v.is_const and v.ty is not None:
            op = type(node.op)
            return const_val({ast.USub: lambda x: -x, ast.UAdd: lambda x: +x, ast.Not: lambda x: not x, ast.Invert: lambda x: ~x}[op](v.const))
        if isinstance(node.op, ast.Not):
            if v.ty is None:
                return const_val(not (v.const if v.is_const else v.obj))
            return Val(f"(!({v.code}))", T.bool_)
        if isinstance(node.op, ast.USub):
            ty = T.promote(v.ty, v.ty)
            return Val(f"(-({self.as_float(v)}))", ty)
        if isinstance(node.op, ast.Invert):
            return Val(f"(~({v.code}))", T.promote(v.ty, v.ty))
        return v

    def binop(self, op: str, l: Val, r: Val, node=None) -> Val:
        if l.is_const and r.is_const:
            try:
                res = _pyfold(op, l.const, r.const)
            except Exception as e:      # noqa: BLE001
                raise self.err(f"constant expression failed: {e}", node)
            return self.wrap(res, node)
        if l.ty is None or r.ty is None:
            raise self.err(f"operator '{op}' on a compile-time-only object", node)
        lp, rp = isinstance(l.ty, (T.Pointer, T.Array)), isinstance(r.ty, (T.Pointer, T.Array))
        if lp or rp:
            if op in ("+", "-") and lp and not rp:
                return Val(f"({l.code} {op} ({r.code}))", T.Pointer(l.ty.elem))
            if op == "+" and rp and not lp:
                return Val(f"({r.code} + ({l.code}))", T.Pointer(r.ty.elem))
            if op == "-" and lp and rp:
                return Val(f"((long long)({l.code} - {r.code}))", T.i64)
            if op in _CMP.values():
                return Val(f"({l.code} {op} {r.code})", T.bool_)
            raise self.err(f"operator '{op}' is not defined for pointers", node)
        if not isinstance(l.ty, T.Scalar) or not isinstance(r.ty, T.Scalar):
            raise self.err(f"operator '{op}' on {l.ty} and {r.ty}", node)
        # a literal adopts the type of the other operand (``tid + 1`` stays int, ``x * 2`` stays unsigned)
        lt, rt = l.ty, r.ty
        if l.is_const and not r.is_const and (rt.is_float or not isinstance(l.const, float)):
            lt = T.promote(rt, rt)
            l = Val(lt.literal(l.const), lt, l.const)
        if r.is_const and not l.is_const and (lt.is_float or not isinstance(r.const, float)):
            rt = T.promote(lt, lt)
            r = Val(rt.literal(r.const), rt, r.const)
        ty = T.promote(lt, rt)
        a, b = self.as_float(l), self.as_float(r)
        if op in _CMP.values():
            return Val(f"({a} {op} {b})", T.bool_)
        if op == "/":
            if ty.is_int:
                return Val(f"((float)({a}) / (float)({b}))", T.f32)
            return Val(f"({a} / {b})", ty)
        if op == "//":
            if ty.is_int:
                return Val(f"({a} / {b})", ty)        # C truncation: equal to Python floor for non-negative operands
            return Val(f"floorf({a} / {b})", ty)
        if op == "%":
            if ty.is_int:
                return Val(f"({a} % {b})", ty)
            return Val(f"fmodf({a}, {b})", ty)
        if op == "**":
            if r.is_const and r.const == 2:
                return Val(f"(({a}) * ({a}))", ty)
            return Val(f"powf({a}, {b})", T.f32)
        if op in ("<<", ">>"):
            if not ty.is_int:
                raise self.err("shift of a floating-point value", node)
            return Val(f"({a} {op} {b})", T.promote(lt, lt))
        if op in ("&", "|", "^"):
            if lt.kind == "b" and rt.kind == "b":
                return Val(f"({a} {op} {b})", T.bool_)
            if not ty.is_int:
                raise self.err("bitwise operator on a floating-point value", node)
        return Val(f"({a} {op} {b})", ty)

    def e_BinOp(self, node):
        return self.binop(_BINOP[type(node.op)], self.expr(node.left), self.expr(node.right), node)

    def e_Compare(self, node):
        left = self.expr(node.left)
        parts: List[Val] = []
        for op, comp in zip(node.ops, node.comparators):
            right = self.expr(comp)
            if type(op) in _CMP:
                parts.append(self.binop(_CMP[type(op)], left, right, node))
            else:       # is / is not / in / not in: compile-time only
                lo = left.const if left.is_const else left.obj
                ro = right.const if right.is_const else right.obj
                if left.ty is not None and not left.is_const or right.ty is not None and not right.is_const:
                    raise self.err("'is' / 'in' need compile-time operands", node)
                res = {ast.Is: lambda: lo is ro, ast.IsNot: lambda: lo is not ro, ast.In: lambda: lo in ro, ast.NotIn: lambda: lo not in ro}[type(op)]()
                parts.append(const_val(res))
            left = right
        return self._bool_join("&&", parts)

    def _bool_join(self, op: str, parts: List[Val]) -> Val:
        keep: List[Val] = []
        for p in parts:
            if p.is_const or p.ty is None:
                truth = bool(p.const if p.is_const else p.obj)
                if op == "&&" and not truth:
                    return const_val(False)
                if op == "||" and truth:
                    return const_val(True)
                continue
            keep.append(p)
        if not keep:
            return const_val(op == "&&")
        if len(keep) == 1:
            v = keep[0]
            return v if v.ty is T.bool_ else Val(f"(({v.code}) != 0)", T.bool_)
        return Val("(" + f" {op} ".join(v.code for v in keep) + ")", T.bool_)

    def e_BoolOp(self, node):
        return self._bool_join("&&" if isinstance(node.op, ast.And) else "||", [self.expr(v) for v in node.values])

    def e_IfExp(self, node):
        c = self.expr(node.test)
        if c.is_const or c.ty is None:
            return self.expr(node.body if (c.const if c.is_const else c.obj) else node.orelse)
        a, b = self.expr(node.body), self.expr(node.orelse)
        if isinstance(a.ty, T.Scalar) and isinstance(b.ty, T.Scalar):
            ty = T.promote(a.ty, b.ty) if not (a.ty.name == b.ty.name) else a.ty
            if a.is_const and not b.is_const:
                ty = b.ty
            if b.is_const and not a.is_const:
                ty = a.ty
            return Val(f"({c.code} ? {self.cast(a, ty).code} : {self.cast(b, ty).code})", ty)
        return Val(f"({c.code} ? {a.code} : {b.code})", a.ty)

    def _index(self, base: Val, idx_node, node) -> Val:
        if isinstance(base.ty, T.Array):
            idxs = idx_node.elts if isinstance(idx_node, ast.Tuple) else [idx_node]
            if len(idxs) > len(base.ty.shape):
                raise self.err("too many indices", node)
            flat: Optional[Val] = None
            for k, e in enumerate(idxs):
                v = self.expr(e)
                stride = 1
                for s in base.ty.shape[k + 1:]:
                    stride *= s
                term = self.binop("*", v, const_val(stride), node) if stride != 1 else v
                flat = term if flat is None else self.binop("+", flat, term, node)
            if len(idxs) < len(base.ty.shape):      # a row of a multi-dimensional array: pointer to it
                return Val(f"({base.code} + {self.rvalue(flat)})", T.Array(base.ty.elem, base.ty.shape[len(idxs):], base.ty.space))
            return Val(f"{base.code}[{self.rvalue(flat)}]", base.ty.elem, lvalue=True)
        if isinstance(base.ty, T.Pointer):
            i = self.expr(idx_node)
            return Val(f"{base.code}[{self.rvalue(i)}]", base.ty.elem, lvalue=True)
        raise self.err(f"a value of type {base.ty} cannot be indexed", node)

    def e_Subscript(self, node):
        b = self.expr(node.value)
        if b.ty is None:
            o = b.const if b.is_const else b.obj
            i = self.expr(node.slice)
            key = i.const if i.is_const else i.obj
            if not i.is_const and i.ty is not None:
                # a Python list / tuple of numbers indexed at run time: materialised once as a constant lookup table
                if isinstance(o, (list, tuple)) and o and all(isinstance(x, (bool, int, float)) for x in o):
                    return Val(f"{self.const_table(o)}[{i.code}]", self._table_type(o))
                raise self.err("indexing a compile-time sequence at run time needs a flat list / tuple of numbers", node)
            try:
                return self.wrap(o[key], node)
            except Exception as e:      # noqa: BLE001
                raise self.err(f"compile-time subscript failed: {e}", node)
        return self._index(b, node.slice, node)

    def e_Call(self, node):
        return self.call(node, None)

    # ------------------------------------------------------------------------------------------------------
    def call(self, node, target: Optional[str]) -> Val:
        f = self.expr(node.func)
        fo = f.obj if f.obj is not None else (f.const if f.is_const else None)
        if fo is None:
            raise self.err("call of a run-time value", node)
        args = [self.expr(a) for a in node.args]
        kwargs = {k.arg: self.expr(k.value) for k in node.keywords}
        if isinstance(fo, T.Scalar):
            if len(args) != 1:
                raise self.err(f"{fo.name}(x) takes one argument", node)
            return self.cast(args[0], fo)
        if isinstance(fo, ll.Intrinsic):
            if fo.declares:
                if target is None:
                    raise self.err(f"ll.{fo.name}(...) must be assigned to a name", node)
                return fo.emit(self, args, kwargs, node, target)
            return fo.emit(self, args, kwargs, node)
        if fo in (ll.unroll, ll.static_range, range):
            if all(a.is_const for a in args):
                return Val("", None, range(*[a.const for a in args]))
            raise self.err("range(...) with run-time bounds is only valid as a for-loop iterator", node)
        if fo in (min, max) and not all(a.is_const for a in args):
            ty = args[0].ty
            for a in args[1:]:
                ty = T.promote(ty, a.ty) if not (a.is_const and not isinstance(a.const, float)) else ty
            cur = self.cast(args[0], ty).code
            for a in args[1:]:
                cur = f"{fo.__name__}({cur}, {self.cast(a, ty).code})"
            return Val(cur, ty)
        if fo is abs and not args[0].is_const:
            return Val(f"fabsf({self.as_float(args[0])})", T.f32) if args[0].ty.is_float else Val(f"abs({args[0].code})", args[0].ty)
        if fo in (int, float, bool) and not args[0].is_const:
            return self.cast(args[0], {int: T.i32, float: T.f32, bool: T.bool_}[fo])
        if fo is len and isinstance(args[0].ty, T.Array):
            return const_val(args[0].ty.shape[0])
        if isinstance(fo, DeviceFunction) or (isinstance(fo, pytypes.FunctionType) and not all(a.ty is None or a.is_const for a in args)):
            return self.call_device(fo, args, kwargs, node)
        if callable(fo):
            if not all(a.is_const or a.ty is None for a in list(args) + list(kwargs.values())):
                raise self.err(f"{getattr(fo, '__name__', fo)!r} is not a DSL function and was called with run-time arguments", node)
            un = lambda v: v.const if v.is_const else v.obj      # noqa: E731
            try:
                return self.wrap(fo(*[un(a) for a in args], **{k: un(v) for k, v in kwargs.items()}), node)
            except Exception as e:      # noqa: BLE001
                raise self.err(f"compile-time call failed: {e}", node)
        raise self.err(f"{fo!r} is not callable", node)

    def call_device(self, fo, args: List[Val], kwargs: Dict[str, Val], node) -> Val:
        pyfn = fo.fn if isinstance(fo, DeviceFunction) else fo
        sig = inspect.signature(pyfn)
        try:
            bound = sig.bind(*args, **kwargs)
        except TypeError as e:
            raise self.err(f"{pyfn.__name__}: {e}", node)
        ordered: List[Optional[Val]] = [bound.arguments.get(n) for n in sig.parameters]

        def keyof(v: Optional[Val]):
            if v is None:
                return ("default",)
            if v.ty is None:
                o = v.const if v.is_const else v.obj
                try:
                    hash(o)
                    return ("c", o)
                except TypeError:
                    return ("c", repr(o))
            ann_const = False
            return ("t", getattr(v.ty, "cname", ""), getattr(v.ty, "shape", None), ann_const)
        # constexpr-annotated parameters specialise on the value
        anns = {n: p.annotation for n, p in sig.parameters.items()}
        keyparts = []
        for (n, _), v in zip(sig.parameters.items(), ordered):
            a = anns[n]
            is_ce = a is T.constexpr or (isinstance(a, str) and "constexpr" in a)
            if v is not None and v.is_const and is_ce:
                keyparts.append(("c", v.const))
            else:
                keyparts.append(keyof(v))
        key = (pyfn, tuple(keyparts))
        hit = self.module.device_keys.get(key)
        if hit is None:
            if key in self.module.in_progress:
                raise self.err(f"recursive call of '{pyfn.__name__}'", node)
            self.module.in_progress.add(key)
            cname = self.module.unique("lk_" + pyfn.__name__)
            # strip compile-time values of non-constexpr parameters down to typed run-time operands
            sub_args = []
            for (n, _), v in zip(sig.parameters.items(), ordered):
                a = anns[n]
                is_ce = a is T.constexpr or (isinstance(a, str) and "constexpr" in a)
                if v is None or is_ce or v.ty is None:
                    sub_args.append(v)
                else:
                    sub_args.append(Val(v.code, v.ty))
            sub = FnCompiler(self.module, pyfn, False, sub_args, cname)
            src = sub.compile_function()
            self.module.device_src.append(src)
            hit = (cname, sub.ret_type or T.void, [p for p in sub.params])
            self.module.device_keys[key] = hit
            self.module.in_progress.discard(key)
        cname, ret, params = hit
        # run-time operands in parameter order (compile-time parameters are baked in)
        rt = []
        pi = 0
        for (n, _), v in zip(sig.parameters.items(), ordered):
            a = anns[n]
            is_ce = a is T.constexpr or (isinstance(a, str) and "constexpr" in a)
            if v is None or is_ce or v.ty is None:
                continue
            pty = params[pi][1]
            pi += 1
            rt.append(self.cast(v, pty).code if isinstance(pty, T.Scalar) else v.code)
        return Val(f"{cname}({', '.join(rt)})", ret)

    # ------------------------------------------------------------------------------------------------------
    # statements
    # ------------------------------------------------------------------------------------------------------
    def stmts(self, body):
        for s in body:
            m = getattr(self, "s_" + type(s).__name__, None)
            if m is None:
                raise self.err(f"unsupported statement: {type(s).__name__}", s)
            m(s)

    def s_Pass(self, node):
        pass

    def s_Expr(self, node):
        if isinstance(node.value, ast.Constant) and isinstance(node.value.value, str):
            return      # docstring
        v = self.call(node.value, None) if isinstance(node.value, ast.Call) else self.expr(node.value)
        if v.code:
            self.emit(v.code + ";")

    def assign_name(self, name: str, v: Val, ann=None, node=None):
        if isinstance(ann, T.ConstExpr) or (v.ty is None and ann is None):
            if v.ty is not None and not v.is_const:
                raise self.err(f"'{name}' is declared constexpr but its value is not known at compile time", node)
            if name in self.vars:
                raise self.err(f"'{name}' is a run-time variable; it cannot be rebound to a compile-time object", node)
            self.cenv[name] = v.const if v.is_const else v.obj
            return
        if name in self.cenv and name not in self.vars:
            if v.is_const and self.loops and self.loops[-1] == "static":
                self.cenv[name] = v.const
                return
            raise self.err(f"'{name}' is a compile-time constant here; annotate the first assignment with a DSL type to make it a variable", node)
        if isinstance(v.ty, T.Array) and name not in self.vars:        # alias of an array (row view): pointer variable
            self.declare_scalar(name, T.Pointer(v.ty.elem))
            self.vars[name] = v.ty
            self.emit(f"{self.cident(name)} = {v.code};")
            return
        if name not in self.vars:
            ty = ann if isinstance(ann, T.Type) else v.ty
            if ty is None:
                raise self.err(f"cannot infer a type for '{name}'", node)
            if isinstance(ty, T.Scalar) and ty.kind == "i" and ty.bits < 32 and ann is None:
                ty = T.i32
            self.declare_scalar(name, ty)
        ty = self.vars[name]
        if isinstance(ty, T.Array):
            raise self.err(f"array '{name}' cannot be reassigned", node)
        if v.ty is not None and (isinstance(ty, T.Struct) != isinstance(v.ty, T.Struct)
                                 or (isinstance(ty, T.Struct) and ty is not v.ty)
                                 or (isinstance(ty, T.Pointer) != isinstance(v.ty, (T.Pointer, T.Array)) and not v.is_const)):
            raise self.err(f"'{name}' is a {ty.cname} (the type of its first assignment in this function) and cannot hold a "
                           f"{v.ty.cname}: locals have one type per function -- use another name", node)
        code = self.cast(v, ty).code if isinstance(ty, (T.Scalar, T.Pointer)) else self.rvalue(v)
        self.emit(f"{self.cident(name)} = {code};")

    def assign_target(self, tgt, v: Val, node):
        if isinstance(tgt, ast.Name):
            return self.assign_name(tgt.id, v, None, node)
        if isinstance(tgt, (ast.Subscript, ast.Attribute)):
            lv = self.expr(tgt)
            if not lv.lvalue:
                raise self.err("cannot assign to this expression", node)
            code = self.cast(v, lv.ty).code if isinstance(lv.ty, (T.Scalar, T.Pointer)) else self.rvalue(v)
            self.emit(f"{lv.code} = {code};")
            return
        raise self.err("unsupported assignment target", node)

    def s_Assign(self, node):
        if len(node.targets) != 1:
            raise self.err("chained assignment is not supported", node)
        tgt = node.targets[0]
        if isinstance(tgt, ast.Tuple):
            if not isinstance(node.value, ast.Tuple) or len(node.value.elts) != len(tgt.elts):
                raise self.err("tuple assignment needs a tuple of the same length on the right", node)
            vals = [self.expr(e) for e in node.value.elts]
            tmps = []
            for v in vals:      # evaluate everything first (a, b = b, a)
                if v.ty is not None and not v.is_const:
                    t = self.fresh()
                    self.decls.append(f"{v.ty.cname} {t};")
                    self.emit(f"{t} = {v.code};")
                    tmps.append(Val(t, v.ty))
                else:
                    tmps.append(v)
            for t, v in zip(tgt.elts, tmps):
                self.assign_target(t, v, node)
            return
        if isinstance(node.value, ast.Call) and isinstance(tgt, ast.Name):
            v = self.call(node.value, tgt.id)
            if isinstance(v.ty, T.Array) and v.code == tgt.id:
                return      # a declaration
            return self.assign_name(tgt.id, v, None, node)
        self.assign_target(tgt, self.expr(node.value), node)

    def s_AnnAssign(self, node):
        ann = self._eval_annotation(node.annotation)
        if not isinstance(node.target, ast.Name):
            raise self.err("annotated assignment needs a plain name", node)
        if node.value is None:
            if isinstance(ann, T.Type) and node.target.id not in self.vars:
                self.declare_scalar(node.target.id, ann)
            return
        v = self.call(node.value, node.target.id) if isinstance(node.value, ast.Call) else self.expr(node.value)
        if isinstance(v.ty, T.Array) and v.code == node.target.id:
            return
        self.assign_name(node.target.id, v, ann, node)

    def s_AugAssign(self, node):
        cur = self.expr(node.target)
        v = self.binop(_BINOP[type(node.op)], cur, self.expr(node.value), node)
        if isinstance(node.target, ast.Name) and node.target.id in self.cenv and node.target.id not in self.vars:
            return self.assign_name(node.target.id, v, None, node)
        self.assign_target(node.target, v, node)

    def s_If(self, node):
        c = self.expr(node.test)
        if c.is_const or c.ty is None:
            self.stmts(node.body if (c.const if c.is_const else c.obj) else node.orelse)
            return
        self.emit(f"if ({c.code}) {{")
        self.ind += 1
        self.stmts(node.body)
        self.ind -= 1
        if node.orelse:
            if len(node.orelse) == 1 and isinstance(node.orelse[0], ast.If):
                self.emit("} else {")
            else:
                self.emit("} else {")
            self.ind += 1
            self.stmts(node.orelse)
            self.ind -= 1
        self.emit("}")

    def s_While(self, node):
        if node.orelse:
            raise self.err("while/else is not supported", node)
        c = self.expr(node.test)
        if c.is_const and not c.const:
            return
        self.emit(f"while ({'true' if c.is_const else c.code}) {{")
        self.ind += 1
        self.loops.append("rt")
        self.stmts(node.body)
        self.loops.pop()
        self.ind -= 1
        self.emit("}")

    def s_Break(self, node):
        if not self.loops or self.loops[-1] != "rt":
            raise self.err("'break' inside a statically unrolled loop", node)
        self.emit("break;")

    def s_Continue(self, node):
        if not self.loops or self.loops[-1] != "rt":
            raise self.err("'continue' inside a statically unrolled loop", node)
        self.emit("continue;")

    def s_For(self, node):
        if node.orelse:
            raise self.err("for/else is not supported", node)
        if not isinstance(node.target, ast.Name):
            raise self.err("the loop variable must be a plain name", node)
        it = node.iter
        pragma = None
        static = False
        if isinstance(it, ast.Call):
            f = self.expr(it.func)
            if f.obj is ll.unroll:
                pragma = "#pragma unroll" + (f" {self.expr(it.args[1]).const}" if len(it.args) > 1 else "")
                it = it.args[0]
            elif f.obj is ll.static_range:
                static = True
        name = node.target.id
        is_range = isinstance(it, ast.Call) and self.expr(it.func).obj in (range, ll.static_range)
        if is_range:
            a = [self.expr(x) for x in it.args]
            if static or (False):
                if not all(x.is_const for x in a):
                    raise self.err("ll.static_range needs compile-time bounds", node)
                seq = range(*[x.const for x in a])
                return self._static_loop(name, seq, node)
            start, stop, step = (const_val(0), a[0], const_val(1)) if len(a) == 1 else (a[0], a[1], a[2] if len(a) > 2 else const_val(1))
            if not step.is_const:       # a run-time step is taken to be positive (grid-stride loops)
                t = self.fresh("step")
                self.decls.append(f"{T.promote(step.ty, step.ty).cname} {t};")
                self.emit(f"{t} = {step.code};")
                step = Val(t, step.ty)
            ty = T.i32
            for x in (start, stop):
                if not x.is_const:
                    ty = T.promote(ty, x.ty) if x.ty.bits > 32 or x.ty.kind == "u" else ty
            if name not in self.vars:
                if name in self.cenv:
                    del self.cenv[name]
                self.declare_scalar(name, ty)
            c = self.cident(name)
            cmp = ">" if step.is_const and step.const < 0 else "<"
            if pragma:
                self.emit(pragma)
            self.emit(f"for ({c} = {self.cast(start, self.vars[name]).code}; {c} {cmp} {self.cast(stop, self.vars[name]).code}; {c} += {step.code}) {{")
            self.ind += 1
            self.loops.append("rt")
            self.stmts(node.body)
            self.loops.pop()
            self.ind -= 1
            self.emit("}")
            return
        seq = self.expr(it)
        if seq.ty is None and (seq.is_const or seq.obj is not None):
            return self._static_loop(name, seq.const if seq.is_const else seq.obj, node)
        raise self.err("for loops iterate over range(...) or a compile-time sequence", node)

    def _static_loop(self, name, seq, node):
        if name in self.vars:
            raise self.err(f"'{name}' is a run-time variable and cannot be a static loop index", node)
        self.loops.append("static")
        for v in seq:
            self.cenv[name] = v
            self.emit("{")
            self.ind += 1
            self.stmts(node.body)
            self.ind -= 1
            self.emit("}")
        self.loops.pop()
        self.cenv.pop(name, None)       # the index does not outlive the loop (it may become a run-time variable later)

    def s_Return(self, node):
        if node.value is None:
            self.emit("return;")
            return
        if self.is_kernel:
            raise self.err("a kernel cannot return a value", node)
        v = self.expr(node.value)
        if v.ty is None:
            raise self.err("device functions return run-time values (compute constants in plain Python helpers)", node)
        if self.ret_type is None:
            self.ret_type = v.ty if not isinstance(v.ty, T.Array) else T.Pointer(v.ty.elem)
        code = self.cast(v, self.ret_type).code if isinstance(self.ret_type, T.Scalar) else v.code
        self.emit(f"return {code};")

    def s_Assert(self, node):
        c = self.expr(node.test)
        if c.is_const or c.ty is None:
            if not (c.const if c.is_const else c.obj):
                msg = self.expr(node.msg).const if node.msg is not None else ast.unparse(node.test)
                raise self.err(f"static assertion failed: {msg}", node)
            return
        self.emit(f"if (!({c.code})) __trap();")

    # ------------------------------------------------------------------------------------------------------
    def param_decl(self, c: str, ty: T.Type) -> str:
        if isinstance(ty, T.TmaDescriptorType):
            return f"const __grid_constant__ CUtensorMap {c}" if self.is_kernel else f"const CUtensorMap& {c}"
        if isinstance(ty, T.Array):
            return f"{ty.elem.cname}* {c}"
        return f"{ty.cname} {c}"

    def compile_function(self, launch_bounds: Optional[int] = None, min_blocks: Optional[int] = None) -> str:
        self.stmts(self.fdef.body)
        params = ", ".join(self.param_decl(c, t) for c, t in self.params)
        head: List[str] = []
        if self.is_kernel:
            lb = f" __launch_bounds__({launch_bounds}{', ' + str(min_blocks) if min_blocks else ''})" if launch_bounds else ""
            head.append(f'extern "C" __global__ void{lb} {self.cname}({params}) {{')
            if self.uses_dyn:
                head.append("  extern __shared__ __align__(16) uint8_t lk_dyn_raw[];")
                if self.dyn_align > 16:
                    a = self.dyn_align
                    head.append(f"  uint8_t* const lk_dyn = lk_dyn_raw + (({a}u - (td::ptx::smem_u32(lk_dyn_raw) & {a - 1}u)) & {a - 1}u);")
                else:
                    head.append("  uint8_t* const lk_dyn = lk_dyn_raw;")
        else:
            ret = (self.ret_type or T.void).cname
            head.append(f"__device__ __forceinline__ {ret} {self.cname}({params}) {{")
        body = head + ["  " + d for d in self.decls] + self.out + ["}"]
        return "\n".join(body)
