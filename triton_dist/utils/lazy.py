"""Declare-then-materialise symmetric allocation (reference: LazyAllocator / NVSHMEMLazyAllocator,
/root/reference/python/triton_dist/utils.py:1063-1537 -- used by the fused EP layer to size the heap).

Tensors are *declared* first (shape/dtype/name); ``total_bytes()``/``breakdown()`` answer sizing questions
without touching memory; ``materialize()`` performs the collective allocations in declaration order."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch


@dataclass
class LazyTensorSpec:
    """What a lazy tensor will be once materialised (reference: utils.py:1064-1090)."""
    name: str
    shape: List[int]
    dtype: torch.dtype
    fill_value: Optional[float] = None

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= int(s)
        return n

    @property
    def nbytes(self) -> int:
        return self.numel * torch.empty(0, dtype=self.dtype).element_size()


class LazyTensor:
    def __init__(self, name: str, shape: Tuple[int, ...], dtype: torch.dtype):
        self.name, self.shape, self.dtype = name, tuple(int(s) for s in shape), dtype
        self._tensor: Optional[torch.Tensor] = None

    @property
    def nbytes(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n * torch.empty(0, dtype=self.dtype).element_size()

    @property
    def materialized(self) -> bool:
        return self._tensor is not None

    def get(self) -> torch.Tensor:
        if self._tensor is None:
            raise RuntimeError(f"LazyTensor '{self.name}' used before LazyAllocator.materialize()")
        return self._tensor

    @property
    def spec(self) -> LazyTensorSpec:
        return LazyTensorSpec(self.name, list(self.shape), self.dtype, getattr(self, "fill_value", None))

    is_materialized = materialized

    def get_underlying_tensor(self) -> Optional[torch.Tensor]:
        return self._tensor

    def __repr__(self):
        return f"LazyTensor({self.name}, {self.shape}, {self.dtype}, materialized={self.materialized})"


class LazyAllocator:
    ALIGN = 1024

    def __init__(self, symmetric: bool = True, device=None):
        self.symmetric = symmetric
        self.device = device
        self._decls: List[LazyTensor] = []

    def declare(self, name: str, shape, dtype: torch.dtype) -> LazyTensor:
        lt = LazyTensor(name, tuple(shape), dtype)
        self._decls.append(lt)
        return lt

    # reference spelling
    def create_tensor(self, shape, dtype, name: Optional[str] = None) -> LazyTensor:
        return self.declare(name or f"t{len(self._decls)}", shape, dtype)

    def total_bytes(self) -> int:
        return sum((d.nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN for d in self._decls)

    def breakdown(self) -> Dict[str, int]:
        return {d.name: d.nbytes for d in self._decls}

    def materialize(self):
        from . import get_heap, current_device
        for d in self._decls:
            if d.materialized:
                continue
            if self.symmetric:
                d._tensor = get_heap().tensor(d.shape, d.dtype)
            else:
                d._tensor = torch.zeros(d.shape, dtype=d.dtype, device=self.device or current_device())
        return self

    sync = materialize

    def free(self):
        from . import get_heap
        for d in reversed(self._decls):
            if d.materialized and self.symmetric:
                get_heap().free_tensor(d._tensor)
            d._tensor = None


class NVSHMEMLazyAllocator(LazyAllocator):
    def __init__(self):
        super().__init__(symmetric=True)


def get_underlying_tensor(t):
    """A LazyTensor's tensor (None before materialisation); plain tensors pass through."""
    return t.get_underlying_tensor() if isinstance(t, LazyTensor) else t


def nvshmem_free_lazy_tensor(t):
    """Collective free of one materialised lazy tensor."""
    from . import nvshmem_free_tensor_sync
    if isinstance(t, LazyTensor):
        if t._tensor is not None:
            nvshmem_free_tensor_sync(t._tensor)
            t._tensor = None
    elif t is not None:
        nvshmem_free_tensor_sync(t)
