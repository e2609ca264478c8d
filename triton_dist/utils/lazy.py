"""Declare-then-materialise symmetric allocation (reference: LazyAllocator / NVSHMEMLazyAllocator,
/root/reference/python/triton_dist/utils.py:1063-1537 -- used by the fused EP layer to size the heap).

Tensors are *declared* first (shape/dtype/name); ``total_bytes()``/``breakdown()`` answer sizing questions
without touching memory; ``materialize()`` performs the collective allocations in declaration order."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


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
