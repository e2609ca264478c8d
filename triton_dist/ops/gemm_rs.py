"""GEMM + ReduceScatter in one kernel:  ``C[M/W, N] = reduce_scatter(A[M, K/W] @ B[K/W, N])``.

Reference: ``create_gemm_rs_context`` / ``gemm_rs`` (/root/reference/python/triton_dist/kernels/nvidia/
gemm_reduce_scatter.py:71, :727-741) = Triton GEMM writing a symmetric buffer + per-segment counters ->
copy-engine scatter -> barrier -> separate ``ring_reduce`` kernel (reduce_scatter.py:551-707).

Here (csrc/gemm_sm100.cuh, mode kRS) the reduce-scatter is a ring fused into the tcgen05 epilogue: rank ``o-1``
computes owner ``o``'s tile first and pushes it (coalesced 16-byte NVLink stores) into rank ``o-2``'s staging
buffer; every rank adds its TMEM accumulator to the running partial it received and forwards it; rank ``o`` adds
the last contribution and writes the final rows.  All SMs run GEMM tiles, there is no reduction pass and no
barrier: staging is double buffered by call parity, per-tile flags carry monotone phase numbers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _C
from .. import utils as U
from .gemm import GemmConfig, attach_splitk, fill_common, gemm
from .ag_gemm import _as_nk


@dataclass
class GEMMReduceScatterTensorParallelContext:
    max_M: int
    N: int
    rank: int
    world_size: int
    local_world_size: int
    output_dtype: torch.dtype
    stage: torch.Tensor = None      # symmetric [2, max_M, N] (16-bit, or fp32 when fp32_ring)
    fp32_ring: bool = False         # running partial sums travel in fp32: one rounding at the owner instead of W-1
    flags: torch.Tensor = None      # symmetric int32 [2, max_tiles]
    phase: torch.Tensor = None      # local int32 [4]
    scratch: object = None          # AllReduceContext of the ragged-M path (GEMM -> NVLS reduce-scatter), lazily allocated
    host_phase: int = 0

    def finalize(self):
        heap = U.get_heap()
        if self.scratch is not None:
            self.scratch.finalize()
        for t in (self.stage, self.flags):
            if t is not None:
                heap.free_tensor(t)
        self.stage = self.flags = self.scratch = None


def create_gemm_rs_context(max_M: int, N: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                           local_world_size: Optional[int] = None, output_dtype: torch.dtype = torch.bfloat16,
                           rs_stream=None, reduce_st: bool = False, fp32_ring: Optional[bool] = None,
                           **ref_hints) -> GEMMReduceScatterTensorParallelContext:
    """``fp32_ring`` (default: env ``TD_RS_FP32``, off): stage the ring's running partial sums in fp32.  The bf16 ring
    rounds the partial at every hop (W-1 roundings); fp32 staging rounds once at the owner like an fp32 reduce, at twice the
    NVLink bytes (hidden behind the mainloop for K/W >= 2048, see profiles/gemm_rs_numerics.md)."""
    U.accept_ref_hints("create_gemm_rs_context", ref_hints, ())
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    ctx = GEMMReduceScatterTensorParallelContext(max_M, N, rank, world_size, local_world_size or world_size, output_dtype)
    ctx.fp32_ring = U.get_bool_env("TD_RS_FP32", False) if fp32_ring is None else bool(fp32_ring)
    ctx.stage = heap.tensor((2, max_M, N), torch.float32 if ctx.fp32_ring else output_dtype)
    max_tiles = max(((max_M + 127) // 128) * ((N + 31) // 32), world_size, 8)
    ctx.flags = heap.tensor((2, max_tiles), torch.int32)
    ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


def default_rs_config(M: int, N: int, K: int, world: int) -> GemmConfig:
    """One L2 band per owner block (``group_m`` = m-tiles per owner): ring order is preserved, B is re-read once
    per owner block instead of once per m-tile (measured: group_m=1 made the 4096x12288x6144 GEMM DRAM bound)."""
    mr = M // max(world, 1)
    if mr % 256 == 0 and N >= 256:
        if U.get_bool_env("TD_RS_BN192", False) and N % 192 == 0:
            # opt-in (not yet timed on hardware): 192-wide tiles trade 10.4 waves of 256x256 for 13.8 shorter ones on 74 CTA pairs
            return GemmConfig(bn=192, cta_group=2, group_m=max(1, mr // 256), use_tma_store=False)
        return GemmConfig(bn=256, cta_group=2, group_m=max(1, mr // 256), use_tma_store=False)
    gm = max(1, mr // 128)
    if mr % 128 == 0 and N >= 256:
        return GemmConfig(bn=256, cta_group=1, group_m=gm, use_tma_store=False)
    if mr % 128 == 0:
        return GemmConfig(bn=128 if N >= 128 else 64, cta_group=1, group_m=gm, use_tma_store=False)
    return GemmConfig(bn=128, cta_group=1, group_m=1, use_tma_store=False)


def gemm_rs(A: torch.Tensor, B: torch.Tensor, ctx: GEMMReduceScatterTensorParallelContext,
            gemm_config: Optional[GemmConfig] = None, persistent: bool = True, fuse_scatter: bool = True,
            reduce_st: bool = False, out: Optional[torch.Tensor] = None, straggler_option=None, profiler=None,
            skip_wait: bool = False, scale_a=None, scale_b=None, **ref_hints) -> torch.Tensor:
    """A: ``[M, K/W]``, B: ``[K/W, N]`` (``.t()`` view of a ``[N, K/W]`` weight) -> ``[M/W, N]``.
    ``skip_wait`` runs the GEMM-only twin: same tiles, same epilogue and pushes, but nobody waits for the partial of the
    previous rank (numerically meaningless; measures the exposed communication as fused - twin).
    int8 / float8_e4m3fn ``A`` and ``B`` (the reference's per-tensor fp8 / int8 gemm_rs dtypes, test_gemm_rs.py:130-145) run on
    tcgen05 ``kind::i8`` / ``kind::f8f6f4`` with ``scale_a`` / ``scale_b`` applied in the epilogue BEFORE the ring (the running
    partial sums travel dequantised, in bf16 or fp32)."""
    U.accept_ref_hints("gemm_rs", ref_hints, ())
    W = ctx.world_size
    M, K = A.shape
    Bnk = _as_nk(B)
    N = Bnk.shape[0]
    assert M % W == 0 and M <= ctx.max_M and N == ctx.N and Bnk.shape[1] == K
    Mr = M // W
    quant = A.dtype not in (torch.bfloat16, torch.float16, torch.float32)
    if quant:
        from .gemm import _Q8_CODE, _scale_vec, gemm_scaled
    if not A.is_cuda:
        if quant:
            part_in = (A.float() * (1.0 if scale_a is None else _scale_vec(scale_a, M, A.device)[:, None])).to(ctx.output_dtype)
            Bq = (Bnk.float() * (1.0 if scale_b is None else _scale_vec(scale_b, N, A.device)[:, None])).to(ctx.output_dtype)
            return _gemm_rs_host(part_in, Bq, ctx, out)
        return _gemm_rs_host(A, Bnk, ctx, out)
    if out is None:
        out = torch.empty((Mr, N), dtype=ctx.output_dtype if quant else A.dtype, device=A.device)
    cfg = gemm_config or default_rs_config(M, N, K, W)
    if W == 1:
        if quant:
            return gemm_scaled(A, Bnk, scale_a, scale_b, out=out)
        return gemm(A, Bnk, out=out, config=GemmConfig(cfg.bn, cfg.cta_group, 8, True, cfg.num_sms, 0))
    if straggler_option and straggler_option[0] == ctx.rank:
        torch.cuda._sleep(int(straggler_option[1]))
    tm = 128 * cfg.cta_group
    if Mr % tm != 0:
        if Mr % 128 == 0:
            cfg = GemmConfig(cfg.bn, 1, 1, False, cfg.num_sms, 0)
            tm = 128
        elif quant:
            raise ValueError("quantised gemm_rs needs (M / world) % 128 == 0")
        else:
            return _gemm_rs_fallback(A, Bnk, ctx, out)
    A = A.contiguous()
    args = _C.GemmArgs()
    args.mode = 2
    fill_common(args, M, A.data_ptr(), A.stride(0), Bnk, out.data_ptr(), Mr, out.stride(0), M, N, K,
                GemmConfig(cfg.bn, cfg.cta_group, cfg.group_m, False, cfg.num_sms, 0), quant or A.dtype == torch.bfloat16)
    if quant:
        assert A.dtype == Bnk.dtype and A.dtype in _Q8_CODE and K % 128 == 0 and out.dtype == torch.bfloat16
        args.is_bf16 = _Q8_CODE[A.dtype]
        sa, sb = _scale_vec(scale_a, M, A.device), _scale_vec(scale_b, N, A.device)
        args.scale_a, args.scale_b = (sa.data_ptr() if sa is not None else None), (sb.data_ptr() if sb is not None else None)
    # first the tiles owned by rank+1 (the chain for owner o starts at rank o-1 and ends at o)
    args.m_rot = (((ctx.rank + 1) % W) * Mr) // tm
    r, w, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = r, w, base, stride, mc
    args.phase = ctx.phase.data_ptr()
    args.rs_rows_per_rank = Mr
    args.rs_stage, args.rs_stage_buf_bytes = ctx.stage.data_ptr(), ctx.max_M * N * ctx.stage.element_size()
    args.rs_flags, args.rs_out, args.rs_ldo = ctx.flags.data_ptr(), out.data_ptr(), out.stride(0)
    args.rs_fp32, args.rs_skip_wait = int(ctx.fp32_ring), int(skip_wait)
    attach_splitk(args, A.device)
    if profiler is not None:
        profiler.attach(args)
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
             "td_gemm_launch(rs)")
    ctx.host_phase += 1
    return out


def _gemm_rs_fallback(A, Bnk, ctx, out):
    """Shapes the ring cannot tile (M/W not a multiple of 128 -- every small-batch decode step): two kernels, no host
    barrier, no eager math.  The tcgen05 GEMM writes its partial product straight into the reduce-scatter staging half
    that the DEVICE-resident call counter selects (CUDA-graph replay safe), then the NVLS pull-reduce kernel
    (``multimem.ld_reduce`` through the switch; P2P loads without NVLS) gives every rank its rows."""
    from . import comm
    M, N = A.shape[0], Bnk.shape[0]
    if ctx.scratch is None:
        # collective lazy allocation: every rank reaches its first ragged call together
        ctx.scratch = comm.create_allreduce_ctx(ctx.max_M * ctx.N * torch.empty(0, dtype=ctx.output_dtype).element_size(),
                                                ctx.rank, ctx.world_size, ctx.local_world_size)
    ar = ctx.scratch
    nbytes = M * N * A.element_size()
    stage0 = ar.stage[:nbytes].view(A.dtype).view(M, N)
    gemm(A, Bnk, out=stage0, out_parity=(ar.phase, ar.workspace_nbytes))
    return comm.reduce_scatter(stage0, ar, output=out, device_parity_input=True)


# ------------------------------------------------------------------------------------------------------------
# emulation (no GPU): the same ring, flags and double buffering on the shared-memory heap
# ------------------------------------------------------------------------------------------------------------
def _gemm_rs_host(A, Bnk, ctx, out):
    import ctypes
    heap = U.get_heap()
    lib = _C.host_lib()
    W, me = ctx.world_size, ctx.rank
    M, K = A.shape
    N = Bnk.shape[0]
    Mr = M // W
    ctx.host_phase += 1
    ph = ctx.host_phase
    par = ph & 1
    timeout = U.get_int_env("TD_HOST_TIMEOUT_US", 60_000_000)
    if out is None:
        out = torch.empty((Mr, N), dtype=A.dtype)
    bt = Bnk.float().t()
    stage = ctx.stage[par]
    prev = (me - 1 + W) % W
    prev_stage = heap.peer_view(ctx.stage, prev)[par]
    for step in range(W):
        owner = (me + 1 + step) % W              # step 0: rank+1's rows ... step W-1: my own rows
        rows = slice(owner * Mr, (owner + 1) * Mr)
        part = A[rows].float() @ bt
        if step > 0:
            # running partial pushed by rank+1 (which computed these rows one step earlier)
            if lib.tdh_wait32(ctypes.c_void_p(ctx.flags.data_ptr() + 4 * (par * ctx.flags.shape[1] + owner)), ph, 1, timeout):
                raise TimeoutError(f"gemm_rs: partial for owner {owner} never arrived (phase {ph})")
            part = part + stage[rows].float()
        if step == W - 1:
            out.copy_(part.to(A.dtype))
        else:
            prev_stage[rows].copy_(part.to(A.dtype))
            flag = heap.peer_ptr(ctx.flags.data_ptr() + 4 * (par * ctx.flags.shape[1] + owner), prev)
            lib.tdh_notify32(ctypes.c_void_p(flag), ph, 1)
    return out


def gemm_rs_mxfp8(a, b, ctx: GEMMReduceScatterTensorParallelContext, out: Optional[torch.Tensor] = None,
                  gemm_config: Optional[GemmConfig] = None) -> torch.Tensor:
    """Block-scaled fp8 GEMM + ring ReduceScatter (BASELINE config #3).  ``a``/``b``: :class:`triton_dist.ops.fp8.MXFP8Tensor`
    of this rank's K-shard (``[M, K/W]`` and ``[N, K/W]``); partial sums travel in bf16 through the same fused-epilogue ring."""
    from .fp8 import dequantize_mxfp8, fill_fp8, gemm_mxfp8
    W = ctx.world_size
    M, K = a.shape
    N = b.shape[0]
    assert M % W == 0 and M <= ctx.max_M and N == ctx.N
    Mr = M // W
    if not a.q.is_cuda:
        return _gemm_rs_host(dequantize_mxfp8(a).to(ctx.output_dtype), dequantize_mxfp8(b).to(ctx.output_dtype), ctx, out)
    if out is None:
        out = torch.empty((Mr, N), dtype=torch.bfloat16, device=a.q.device)
    if W == 1:
        return gemm_mxfp8(a, b, out=out, config=gemm_config)
    cg = 2 if Mr % 256 == 0 else 1
    assert Mr % (128 * cg) == 0, "fp8 gemm_rs needs (M / world) % 128 == 0"
    cfg = gemm_config or GemmConfig(bn=128, cta_group=cg, group_m=max(1, Mr // (128 * cg)), use_tma_store=False)
    tm = 128 * cfg.cta_group
    args = _C.GemmArgs()
    args.mode = 2
    fill_common(args, M, a.q.data_ptr(), a.q.stride(0), b.q, out.data_ptr(), Mr, out.stride(0), M, N, K,
                GemmConfig(cfg.bn, cfg.cta_group, cfg.group_m, False, cfg.num_sms, 0), True)
    fill_fp8(args, a, b)
    args.m_rot = (((ctx.rank + 1) % W) * Mr) // tm
    r, w, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = r, w, base, stride, mc
    args.phase = ctx.phase.data_ptr()
    args.rs_rows_per_rank = Mr
    args.rs_stage, args.rs_stage_buf_bytes = ctx.stage.data_ptr(), ctx.max_M * N * ctx.stage.element_size()
    args.rs_flags, args.rs_out, args.rs_ldo = ctx.flags.data_ptr(), out.data_ptr(), out.stride(0)
    args.rs_fp32 = int(ctx.fp32_ring)
    attach_splitk(args, a.q.device)
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemm_launch(rs, mxfp8)")
    ctx.host_phase += 1
    return out


# ------------------------------------------------------------------------------------------------------------
# autotuned entry point (reference: gemm_rs is autotuned over {tile config x fuse_scatter x persistent}, gemm_reduce_scatter.py:709-741)
# ------------------------------------------------------------------------------------------------------------
from ..tune import autotune  # noqa: E402


def get_gemm_rs_config_space():
    """Tile configurations of the fused GEMM + ReduceScatter kernel worth timing: 2-CTA 256 / 192 / 128-wide tiles and the 1-CTA tile
    (the only one for ``M / world`` that is a multiple of 128 but not of 256).  The ring runs in the epilogue of every one of them."""
    return [dict(bn=256, cta_group=2), dict(bn=192, cta_group=2), dict(bn=128, cta_group=2), dict(bn=256, cta_group=1), dict(bn=128, cta_group=1)]


def gemm_rs_key_fn(A, B, ctx, **_):
    return f"{tuple(A.shape)}x{tuple(B.shape)}@tp{ctx.world_size}:{A.dtype}"


def gemm_rs_prune_fn(cfg, A, B, ctx, **_):
    """Keep a configuration only if the kernel accepts it for this shape."""
    Mr = A.shape[0] // ctx.world_size
    N = B.shape[1] if B.shape[0] == A.shape[1] else B.shape[0]
    if Mr % (128 * cfg["cta_group"]):
        return False
    return N % cfg["bn"] == 0 or cfg["bn"] != 192


@autotune(get_gemm_rs_config_space(), key_fn=gemm_rs_key_fn, prune_fn=gemm_rs_prune_fn, warmup=3, rep=8)
def gemm_rs_tuned(A: torch.Tensor, B: torch.Tensor, ctx: GEMMReduceScatterTensorParallelContext, out: Optional[torch.Tensor] = None,
                  config: Optional[dict] = None, **kw) -> torch.Tensor:
    """``gemm_rs`` with the tile configuration chosen by the function-level autotuner (max over ranks decides, cached on disk per shape /
    world size / GPU); ``autotune=False`` takes the first configuration."""
    c = config or get_gemm_rs_config_space()[0]
    return gemm_rs(A, B, ctx, gemm_config=GemmConfig(c["bn"], c["cta_group"], 8, True, 0, 0) if A.is_cuda else None, out=out, **kw)

