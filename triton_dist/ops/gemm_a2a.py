"""GEMM fused with an all-to-all of its output columns (Ulysses QKV projection, inference and training).

``out_on_rank_d[src * rows + r, :] = (x_src @ w[d * c:(d + 1) * c].T)[r, :]`` -- every rank multiplies its sequence shard
by the whole weight; column block d of the product belongs to rank d (its heads).  The tcgen05 GEMM's epilogue stores
each tile *directly into the destination rank's receive buffer* over NVLink (mode kAR / scatter flavour of
csrc/gemm_sm100.cuh), counts finished tiles per destination and raises that destination's flag after the last one, so the
exchange is spread over the whole GEMM instead of following it.  A tiny consumer kernel waits for the W flags (and
optionally copies the double-buffered receive area to a stable tensor, which keeps the op CUDA-graph replayable).

Reference: ``ulysses_sp_infer_gemm_a2a_op`` / ``kernel_gemm_a2a_producer_gemm_with_quant_persistent``
(kernels/nvidia/ulysses_sp_infer_gemm_a2a.py:143-289,455) and ``SpUlysessQKVGemmAll2AllKernel``
(sp_ulysess_qkv_gemm_all2all.py:64-196).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _C
from .. import utils as U
from .gemm import _Q8_CODE, GemmConfig, _scale_vec, fill_common, gemm_scaled

_C.register("td_wait_phase_copy", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p])


@dataclass
class GemmA2AContext:
    max_rows: int                 # rows one rank contributes (its sequence shard)
    cols_per_rank: int
    dtype: torch.dtype
    rank: int
    world_size: int
    recv: torch.Tensor = None     # symmetric [2, W * max_rows, cols_per_rank]
    flags: torch.Tensor = None    # symmetric int32 [2, W]
    phase: torch.Tensor = None    # local int32 [4]
    count: torch.Tensor = None    # local int32 [W]
    host_phase: int = 0
    a2a_ctx: object = None        # all-to-all workspace of the two-kernel quantised path (created on first use)

    def finalize(self):
        heap = U.get_heap()
        if self.a2a_ctx is not None:
            self.a2a_ctx.finalize()
            self.a2a_ctx = None
        for t in (self.recv, self.flags):
            if t is not None:
                heap.free_tensor(t)
        self.recv = self.flags = None


def create_gemm_a2a_context(max_rows: int, cols_per_rank: int, dtype: torch.dtype = torch.bfloat16, rank: Optional[int] = None,
                            world_size: Optional[int] = None) -> GemmA2AContext:
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    ctx = GemmA2AContext(max_rows, cols_per_rank, dtype, rank, world_size)
    ctx.recv = heap.tensor((2, world_size * max_rows, cols_per_rank), dtype)
    ctx.flags = heap.tensor((2, world_size), torch.int32)
    ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
    ctx.count = torch.zeros(max(world_size, 4), dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


def default_a2a_config(M: int, cols_per_rank: int) -> GemmConfig:
    for bn in (256, 128, 64, 32):
        if cols_per_rank % bn == 0:
            return GemmConfig(bn=bn, cta_group=2 if (M > 128 and bn >= 128) else 1, group_m=8, use_tma_store=False)
    raise ValueError("cols_per_rank must be a multiple of 32")


def gemm_all_to_all(ctx: GemmA2AContext, x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None,
                    gemm_config: Optional[GemmConfig] = None, scale_a=None, scale_b=None, fused: Optional[bool] = None) -> torch.Tensor:
    """x: [rows, K] (this rank's sequence shard); w: [W * cols_per_rank, K], rows grouped by destination rank.
    Returns ``[W * rows, cols_per_rank]``: block s = rank s's rows projected onto MY columns.

    8-bit operands (int8 / float8_e4m3fn) with ``scale_a`` (per tensor or per row) and ``scale_b`` (per tensor or per output channel):
    the quantised inference flavour of the reference (ulysses_sp_infer_gemm_a2a.py:143-260, int8 x int8 with input / weight scales).
    The dequantisation happens in the tcgen05 epilogue; the result travels as bf16.  ``fused=True`` (or ``TD_GEMM_A2A_Q8_FUSED=1``)
    scatters straight from the epilogue of the 8-bit GEMM (one kernel); the default runs the scaled GEMM and then the all-to-all
    kernel -- both halves are hardware-validated, the fused combination has not run on a GPU yet."""
    W, me, c = ctx.world_size, ctx.rank, ctx.cols_per_rank
    M, K = x.shape
    assert w.shape == (W * c, K) and M <= ctx.max_rows
    q8 = x.dtype in _Q8_CODE
    if not q8 and (scale_a is not None or scale_b is not None):
        raise NotImplementedError("gemm_all_to_all: scale_a / scale_b apply to int8 / float8_e4m3fn operands")
    if not x.is_cuda:
        return _gemm_a2a_host(ctx, x, w, out, scale_a, scale_b)
    odt = torch.bfloat16 if q8 else x.dtype
    assert ctx.dtype == odt, f"the context was created for {ctx.dtype}, the result is {odt}"
    if out is None:
        out = torch.empty((W * M, c), dtype=odt, device=x.device)
    x = x.contiguous()
    if q8:
        if fused is None:
            fused = os.environ.get("TD_GEMM_A2A_Q8_FUSED", "0") == "1"
        if not fused:
            from .all_to_all import all_to_all_single_2d, create_all_to_all_single_2d_context
            if ctx.a2a_ctx is None:          # collective: every rank reaches its first quantised call together
                ctx.a2a_ctx = create_all_to_all_single_2d_context(ctx.max_rows, c, odt)
            y = gemm_scaled(x, w, scale_a, scale_b, config=gemm_config)                     # [M, W * c] bf16
            return all_to_all_single_2d(ctx.a2a_ctx, y.view(M, W, c).transpose(0, 1).reshape(W * M, c), out)
    cfg = gemm_config or default_a2a_config(M, c)
    args = _C.GemmArgs()
    args.mode = 3
    fill_common(args, M, x.data_ptr(), x.stride(0), w, ctx.recv.data_ptr(), M, c, M, W * c, K,
                GemmConfig(cfg.bn, cfg.cta_group, cfg.group_m, False, cfg.num_sms, 0), x.dtype == torch.bfloat16)
    if q8:
        assert K % 128 == 0 and w.dtype == x.dtype
        sa, sb = _scale_vec(scale_a, M, x.device), _scale_vec(scale_b, W * c, x.device)
        args.is_bf16 = _Q8_CODE[x.dtype]
        args.scale_a = sa.data_ptr() if sa is not None else None
        args.scale_b = sb.data_ptr() if sb is not None else None
    r, wd, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = r, wd, base, stride, mc
    args.phase = ctx.phase.data_ptr()
    # my rows start at me * M in every destination's receive area (the consumer views it as [W, M, c])
    args.ag_rows_per_rank, args.ag_copy_local, args.ag_ready = M, c, ctx.count.data_ptr()
    args.rs_rows_per_rank = 0
    esz = out.element_size()
    args.rs_stage, args.rs_stage_buf_bytes = ctx.recv.data_ptr(), W * ctx.max_rows * c * esz
    args.rs_flags, args.rs_out, args.rs_ldo = ctx.flags.data_ptr(), out.data_ptr(), c
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), s), "td_gemm_launch(a2a)")
    _C.check(_C.cuda_lib().td_wait_phase_copy(ctx.flags.data_ptr(), W, ctx.phase.data_ptr(), ctx.recv.data_ptr(),
                                              W * ctx.max_rows * c * esz, out.data_ptr(), W * M * c * esz, s), "td_wait_phase_copy")
    ctx.host_phase += 1
    return out


def _gemm_a2a_host(ctx, x, w, out, scale_a=None, scale_b=None):
    """Emulation: same protocol -- column block d of the local product goes to rank d's receive slot ``me``; one
    release flag per (destination, source) carrying the phase."""
    import ctypes
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, c = ctx.world_size, ctx.rank, ctx.cols_per_rank
    M = x.shape[0]
    ctx.host_phase += 1
    ph, par = ctx.host_phase, ctx.host_phase & 1
    if x.dtype in _Q8_CODE:
        y = gemm_scaled(x, w, scale_a, scale_b)              # CPU branch of the scaled GEMM: fp32 product, scales, bf16
    else:
        y = (x.float() @ w.float().t()).to(x.dtype)
    for j in range(W):
        d = (me + j) % W
        heap.peer_view(ctx.recv, d)[par, me * M:(me + 1) * M] = y[:, d * c:(d + 1) * c]
        lib.tdh_notify32(ctypes.c_void_p(heap.peer_ptr(ctx.flags[par, me:me + 1].data_ptr(), d)), ph, 1)
    if lib.tdh_wait32_n(ctypes.c_void_p(ctx.flags[par].data_ptr()), W, ph, 1, U.get_int_env("TD_HOST_TIMEOUT_US", 60_000_000)):
        raise TimeoutError("gemm_all_to_all: a peer never delivered")
    res = ctx.recv[par, :W * M]
    if out is not None:
        out.copy_(res)
        return out
    return res.clone()
