"""Typed expression values that flow through the code generator."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

from . import types as T

NOCONST = object()


@dataclass
class Val:
    """One lowered expression.  ``code`` is C++ text; ``ty`` its DSL type; ``const`` the Python value when it is known at compile time;
    ``obj`` a Python-level object (a type, an intrinsic, a module, a device function) that is not a run-time value."""
    code: str = ""
    ty: Optional[T.Type] = None
    const: Any = NOCONST
    obj: Any = None
    lvalue: bool = False

    @property
    def is_const(self):
        return self.const is not NOCONST

    @property
    def is_obj(self):
        return self.obj is not None and self.ty is None and self.const is NOCONST


def const_val(v) -> Val:
    if isinstance(v, (bool, int, float)):
        ty = T.type_of_const(v)
        return Val(ty.literal(v), ty, v)
    return Val("", None, v)           # lists, tuples, strings, None: compile-time only


class CompileError(Exception):
    def __init__(self, msg, node=None, fn=None):
        loc = ""
        if node is not None and hasattr(node, "lineno"):
            base = getattr(fn, "first_line", 1) if fn is not None else 1
            name = getattr(fn, "name", "?") if fn is not None else "?"
            loc = f" [{name}:{base + node.lineno - 1}]"
        super().__init__(msg + loc)
