"""GEMM + AllReduce:  ``out[M, N] = all_reduce(A[M, K/W] @ B[N, K/W]^T)``.

Reference: ``create_gemm_ar_context`` / ``gemm_allreduce_op`` / ``low_latency_gemm_allreduce_op``
(/root/reference/python/triton_dist/kernels/nvidia/gemm_allreduce.py:103-127,669-731) and ``GemmARLayer``.

Two realisations:

* ``low_latency_gemm_allreduce_op`` -- ONE kernel (mode kAR of csrc/gemm_sm100.cuh, the counterpart of the reference's
  ``kernel_fused_gemm_allreduce`` :565-604): GEMM CTAs stage each partial tile in the symmetric buffer and raise a
  per-(tile, rank) flag on every rank; comm CTAs of the same grid wait for all W flags of a tile and reduce it with
  ``multimem.ld_reduce`` through the NVSwitch (P2P loads without NVLS) straight into the local output.  Tiles are
  reduced while later tiles are still being computed; double-buffered by the device-resident call counter, so it is
  CUDA-graph replayable and never resets flags.
* ``gemm_allreduce_op`` -- for large M (bandwidth bound): the tcgen05 GEMM's TMA-store epilogue writes the partial
  product directly into the staging buffer of the all-reduce context (zero copy), then the two-shot NVLS all-reduce
  kernel runs on the same stream with its staging copy skipped (two launches, no host sync, graph-capturable).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

import ctypes as C

from .. import _C
from .. import utils as U
from . import comm
from .ag_gemm import _as_nk
from .gemm import GemmConfig, default_config, fill_common, gemm


@dataclass
class GemmARContext:
    max_M: int
    N: int
    dtype: torch.dtype
    rank: int
    world_size: int
    ar_ctx: comm.AllReduceContext = None
    calls: int = 0
    # fused single-kernel path
    stage: torch.Tensor = None       # symmetric [2, max_M, N]
    flags: torch.Tensor = None       # symmetric int32 [2, flag_tiles, W]
    phase: torch.Tensor = None       # local int32 [4] device-resident call counter
    flag_tiles: int = 0
    num_comm_sms: int = 16

    def finalize(self):
        if self.ar_ctx is not None:
            self.ar_ctx.finalize()
            self.ar_ctx = None
        heap = U.get_heap()
        for t in (self.stage, self.flags):
            if t is not None:
                heap.free_tensor(t)
        self.stage = self.flags = None


def create_gemm_ar_context(ar_stream=None, rank: Optional[int] = None, world_size: Optional[int] = None,
                           local_world_size: Optional[int] = None, max_M: int = 0, N: int = 0,
                           dtype: torch.dtype = torch.bfloat16, **ref_hints) -> GemmARContext:
    U.accept_ref_hints("create_gemm_ar_context", ref_hints, ('MIN_BLOCK_SIZE_M', 'MIN_BLOCK_SIZE_N', 'NUM_COMM_SMS', 'TILE_MAP_LEVEL'))
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    ctx = GemmARContext(max_M, N, dtype, rank, world_size)
    nbytes = max_M * N * torch.empty(0, dtype=dtype).element_size()
    ctx.ar_ctx = comm.create_allreduce_ctx(max(nbytes, 1024), rank, world_size, local_world_size or world_size)
    return ctx


def create_ll_gemm_ar_context(ar_stream=None, rank: Optional[int] = None, world_size: Optional[int] = None,
                              local_world_size: Optional[int] = None, max_M: int = 0, N: int = 0,
                              dtype: torch.dtype = torch.bfloat16, NUM_COMM_SMS: int = 16, **ref_hints) -> GemmARContext:
    """Context of the single-kernel GEMM+AllReduce (reference :127, double-buffered ``num_phases=2``)."""
    U.accept_ref_hints("create_ll_gemm_ar_context", ref_hints, ('MIN_BLOCK_SIZE_M', 'MIN_BLOCK_SIZE_N', 'num_phases'))
    ctx = create_gemm_ar_context(ar_stream, rank, world_size, local_world_size, max_M, N, dtype)
    heap = U.get_heap()
    ctx.stage = heap.tensor((2, max_M, N), dtype)
    ctx.flag_tiles = max(((max_M + 127) // 128) * ((N + 31) // 32), 8)
    ctx.flags = heap.tensor((2, ctx.flag_tiles, ctx.world_size), torch.int32)
    ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
    ctx.num_comm_sms = NUM_COMM_SMS
    U.barrier_all_host()
    return ctx


def create_gemm_ar_context_auto(rank: int, world_size: int, max_M: int, N: int, dtype: torch.dtype = torch.bfloat16) -> GemmARContext:
    """What the TP layers use: the single-kernel (low-latency) context for decode-sized batches (max_M <= 256, as the
    reference's GemmARLayer picks its LL kernel, tp_mlp.py:197-199), the two-kernel NVLS context otherwise.
    ``TD_GEMM_AR_FUSED=0`` forces the two-kernel path."""
    make = create_ll_gemm_ar_context if (max_M <= 256 and U.get_bool_env("TD_GEMM_AR_FUSED", True)) else create_gemm_ar_context
    return make(None, rank, world_size, world_size, max_M, N, dtype)


def default_ar_config(M: int, N: int, K: int, n_comm: int = 16, num_sms: int = 148) -> GemmConfig:
    """Small-M decode shapes: narrow tiles so that enough CTAs work on the K-reduction; the comm CTAs take the SMs the
    GEMM has no tile for (a comm CTA reduces its tiles one after the other and each NVLS round trip is ~2 us, so every idle
    SM is worth using: 8xB200, M=128 N=5120: 16 comm CTAs 75 us)."""
    if M > 128 and N >= 256:
        tiles = ((M + 255) // 256) * ((N + 127) // 128) * 2
        return GemmConfig(bn=128, cta_group=2, group_m=8, use_tma_store=False, n_comm_ctas=max(n_comm, min(64, (num_sms - tiles) // 2 * 2)))
    bn = 32 if N < 64 else 64 if N <= 8192 else 128
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    return GemmConfig(bn=bn, cta_group=1, group_m=8, use_tma_store=False, n_comm_ctas=max(n_comm, min(64, num_sms - tiles)))


_FUSED_MAX_M = 32      # rows up to which the one-kernel GEMM + AllReduce is the auto choice (see low_latency_gemm_allreduce_op)


def low_latency_gemm_allreduce_op(ctx: GemmARContext, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
                                  gemm_config: Optional[GemmConfig] = None, straggler_option=None, A_scale=None, B_scale=None,
                                  **ref_hints) -> torch.Tensor:
    """Single fused kernel; falls back to :func:`gemm_allreduce_op` when the context has no fused buffers.  Quantised operands
    (int8 / float8_e4m3fn with ``A_scale`` / ``B_scale``) take the two-kernel path (scaled tcgen05 GEMM -> NVLS all-reduce)."""
    U.accept_ref_hints("low_latency_gemm_allreduce_op", ref_hints, ('copy_to_local', 'USE_MULTIMEM_ST', 'TILE_MAP_LEVEL'))
    if a.dtype not in (torch.bfloat16, torch.float16, torch.float32):
        return gemm_allreduce_op(ctx, a, b, out=out, gemm_config=None, straggler_option=straggler_option, As=A_scale, Bs=B_scale)
    if A_scale is not None or B_scale is not None:
        raise NotImplementedError("low_latency_gemm_allreduce_op: scales apply to int8 / float8_e4m3fn operands")
    w = b if (b.shape[1] == a.shape[1] and b.stride(1) == 1) else _as_nk(b)
    M, K = a.shape
    N = w.shape[0]
    if (not a.is_cuda) or ctx.world_size == 1 or ctx.stage is None:
        return gemm_allreduce_op(ctx, a, w, out, gemm_config, straggler_option=straggler_option)
    if M <= 8 and gemm_config is None:
        # decode rows: the product is pure weight streaming -- the CUDA-core GEMV uses every SM's load bandwidth, a 128-row
        # tensor-core tile would stream the weight through N/64 CTAs only (measured: Qwen3-8B TP2 decode 2.27 ms fused vs
        # 1.59 ms GEMV + one-shot NVLS all-reduce)
        part = gemm(a, w)
        return comm.all_reduce(part, None, ctx.ar_ctx, output=out, straggler_option=straggler_option)
    assert M <= ctx.max_M and N == ctx.N
    if gemm_config is None and M > _FUSED_MAX_M and not U.get_bool_env("TD_GEMM_AR_FORCE_FUSED", False):
        # measured on 8xB200 (profiles/extras_8xB200.json): the single kernel wins at M = 16 (30.0 vs 31.8 us) but loses at M = 128
        # (75 vs 44 us: few comm CTAs reduce many tiles, ~2 us NVLS round trip each) -> above the crossover take the two-kernel path
        # (tcgen05 GEMM into the device-selected staging half, then the two-shot NVLS all-reduce) unless a config is given explicitly
        return gemm_allreduce_op(ctx, a, w, out, None, straggler_option=straggler_option)
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    if straggler_option and straggler_option[0] == ctx.rank:
        torch.cuda._sleep(int(straggler_option[1]))
    cfg = gemm_config or default_ar_config(M, N, K, ctx.num_comm_sms)
    if cfg.n_comm_ctas <= 0:
        cfg = GemmConfig(cfg.bn, cfg.cta_group, cfg.group_m, False, cfg.num_sms, ctx.num_comm_sms)
    a = a.contiguous()
    args = _C.GemmArgs()
    args.mode = 3
    fill_common(args, M, a.data_ptr(), a.stride(0), w, out.data_ptr(), M, out.stride(0), M, N, K,
                GemmConfig(cfg.bn, cfg.cta_group, cfg.group_m, False, cfg.num_sms, cfg.n_comm_ctas), a.dtype == torch.bfloat16)
    r, wd, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = r, wd, base, stride, mc
    args.phase = ctx.phase.data_ptr()
    args.rs_rows_per_rank = ctx.flag_tiles
    args.rs_stage, args.rs_stage_buf_bytes = ctx.stage.data_ptr(), ctx.max_M * N * a.element_size()
    args.rs_flags, args.rs_out, args.rs_ldo = ctx.flags.data_ptr(), out.data_ptr(), out.stride(0)
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
             "td_gemm_launch(ar)")
    ctx.calls += 1
    return out


def gemm_allreduce_op(ctx: GemmARContext, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
                      gemm_config: Optional[GemmConfig] = None, method=None, straggler_option=None, As=None, Bs=None,
                      **ref_hints) -> torch.Tensor:
    """``a``: [M, K/W]; ``b``: weight [N, K/W] (K-major) or its ``.t()`` view -> all-reduced [M, N].
    int8 (or float8_e4m3fn) ``a`` / ``b`` with ``As`` (per-row or per-tensor activation scales) and ``Bs`` (per-output-channel
    or per-tensor weight scales): the reference's quantised GEMM+AllReduce (gemm_allreduce.py:383-447) -- tcgen05 ``kind::i8``
    with the dequantisation in the epilogue, partial sums all-reduced in bf16."""
    U.accept_ref_hints("gemm_allreduce_op", ref_hints, ('copy_to_local', 'USE_MULTIMEM_ST', 'pg'))
    w = b if (b.shape[1] == a.shape[1] and b.stride(1) == 1) else _as_nk(b)
    M, K = a.shape
    N = w.shape[0]
    assert M <= ctx.max_M and N == ctx.N
    quant = a.dtype not in (torch.bfloat16, torch.float16, torch.float32)
    if quant:
        from .gemm import gemm_scaled
        odt = ctx.dtype if ctx.dtype in (torch.bfloat16,) else torch.bfloat16
        if out is None:
            out = torch.empty((M, N), dtype=odt, device=a.device)
        if not a.is_cuda or ctx.world_size == 1:
            part = gemm_scaled(a, w, As, Bs, out_dtype=odt)
            return part if ctx.world_size == 1 else comm.all_reduce(part, comm.AllReduceMethod.OneShot, ctx.ar_ctx, output=out)
        nbytes = M * N * 2
        ctx.calls += 1
        stage0 = ctx.ar_ctx.stage[:nbytes].view(odt).view(M, N)
        gemm_scaled(a, w, As, Bs, out=stage0, config=gemm_config or default_config(M, N, K), out_parity=(ctx.ar_ctx.phase, ctx.ar_ctx.workspace_nbytes))
        comm.all_reduce(stage0, method or comm.get_auto_allreduce_method(nbytes), ctx.ar_ctx, output=out, straggler_option=straggler_option,
                        device_parity_input=True)
        return out
    if As is not None or Bs is not None:
        raise NotImplementedError("gemm_allreduce_op: As / Bs scales apply to int8 / float8_e4m3fn operands")
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
    comm.all_reduce(stage0, method, ctx.ar_ctx, output=out, straggler_option=straggler_option, device_parity_input=True)
    return out



def gemm_op(ctx, a, b, out=None):
    return gemm(a, b if b.stride(1) == 1 else _as_nk(b), out=out)


def allreduce_op(ctx: GemmARContext, x: torch.Tensor, out=None):
    return comm.all_reduce(x, None, ctx.ar_ctx, output=out)
