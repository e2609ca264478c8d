"""GEMM + AllReduce:  ``out[M, N] = all_reduce(A[M, K/W] @ B[N, K/W]^T)``.

Reference: ``create_gemm_ar_context`` / ``gemm_allreduce_op`` / ``low_latency_gemm_allreduce_op``
(/root/reference/python/triton_dist/kernels/nvidia/gemm_allreduce.py:103-127,669-731) and ``GemmARLayer``.

Current realisation: the tcgen05 GEMM's TMA-store epilogue writes the partial product *directly into the
symmetric staging buffer* of the all-reduce context (zero copy), then the NVLS all-reduce kernel
(multimem.ld_reduce / multimem.st, csrc/comm_kernels.cu) runs on the same stream with its staging copy skipped.
Two launches, no host sync, graph-capturable; the message never makes an extra HBM round trip.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .. import utils as U
from . import comm
from .ag_gemm import _as_nk
from .gemm import GemmConfig, default_config, gemm


@dataclass
class GemmARContext:
    max_M: int
    N: int
    dtype: torch.dtype
    rank: int
    world_size: int
    ar_ctx: comm.AllReduceContext = None
    calls: int = 0

    def finalize(self):
        if self.ar_ctx is not None:
            self.ar_ctx.finalize()
            self.ar_ctx = None


def create_gemm_ar_context(ar_stream=None, rank: Optional[int] = None, world_size: Optional[int] = None,
                           local_world_size: Optional[int] = None, max_M: int = 0, N: int = 0,
                           dtype: torch.dtype = torch.bfloat16, **_unused) -> GemmARContext:
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    ctx = GemmARContext(max_M, N, dtype, rank, world_size)
    nbytes = max_M * N * torch.empty(0, dtype=dtype).element_size()
    ctx.ar_ctx = comm.create_allreduce_ctx(max(nbytes, 1024), rank, world_size, local_world_size or world_size)
    return ctx


create_ll_gemm_ar_context = create_gemm_ar_context


def gemm_allreduce_op(ctx: GemmARContext, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
                      gemm_config: Optional[GemmConfig] = None, method=None, straggler_option=None, **_unused) -> torch.Tensor:
    """``a``: [M, K/W]; ``b``: weight [N, K/W] (K-major) or its ``.t()`` view -> all-reduced [M, N]."""
    w = b if (b.shape[1] == a.shape[1] and b.stride(1) == 1) else _as_nk(b)
    M, K = a.shape
    N = w.shape[0]
    assert M <= ctx.max_M and N == ctx.N
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    if not a.is_cuda:
        part = (a.float() @ w.float().t()).to(a.dtype)
        return comm.all_reduce(part, comm.AllReduceMethod.OneShot, ctx.ar_ctx, output=out)
    if ctx.world_size == 1:
        return gemm(a, w, out=out, config=gemm_config)
    nbytes = M * N * a.element_size()
    ctx.calls += 1
    # The all-reduce kernel reduces staging half (phase+1)&1 where `phase` is ITS device-resident call counter.
    # The GEMM reads the same counter on the device and stores into that half, so the pair stays consistent
    # when a captured CUDA graph is replayed.
    ws = ctx.ar_ctx.workspace_nbytes
    stage0 = ctx.ar_ctx.stage[:nbytes].view(a.dtype).view(M, N)
    gemm(a, w, out=stage0, config=gemm_config or default_config(M, N, K), out_parity=(ctx.ar_ctx.phase, ws))
    if method is None:
        method = comm.get_auto_allreduce_method(nbytes)
    comm.all_reduce(stage0, method, ctx.ar_ctx, output=out, straggler_option=straggler_option)
    return out


low_latency_gemm_allreduce_op = gemm_allreduce_op


def gemm_op(ctx, a, b, out=None):
    return gemm(a, b if b.stride(1) == 1 else _as_nk(b), out=out)


def allreduce_op(ctx: GemmARContext, x: torch.Tensor, out=None):
    return comm.all_reduce(x, None, ctx.ar_ctx, output=out)
