"""Remaining entry points of the reference's op namespace, composed from the kernels in this package.

Each is the same contract as the reference function it names; where the reference fuses two steps in one Triton kernel
and we currently run two of our kernels back to back on one stream, the docstring says so.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import utils as U
from . import comm
from . import moe as M
from .all_to_all import (AllToAllContext, all_to_all_single_2d, create_all_to_all_single_2d_context)
from .elementwise import silu_mul
from .gemm import gemm


def _lin(x, w):
    return gemm(x, w) if x.is_cuda else torch.nn.functional.linear(x, w)


# ---- reduce_scatter.py ---------------------------------------------------------------------------------------
def create_reduce_scater_2d_ctx(max_M: int, N: int, rank: int, world_size: int, local_world_size: int, dtype: torch.dtype, **_):
    return comm.create_allreduce_ctx(max_M * N * torch.empty(0, dtype=dtype).element_size(), rank, world_size, local_world_size)


def reduce_scatter_2d_op(input: torch.Tensor, ctx, output: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(reduce_scatter.py:822) scatter + local reduce -> here one pull-reduce kernel (NVLS ld_reduce when available)."""
    return comm.reduce_scatter(input, ctx, output)


def ring_reduce(slabs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(reduce_scatter.py:780) out = sum over the leading (source) dimension, fp32 accumulation."""
    out = torch.empty(slabs.shape[1:], dtype=slabs.dtype, device=slabs.device) if out is None else out
    return comm.reduce_tensor(out, slabs.contiguous())


# ---- swiglu.py ------------------------------------------------------------------------------------------------
from .elementwise import silu_mul_backward as swiglu_backward  # noqa: E402,F401  (one kernel: csrc/elementwise.cu silu_mul_bwd_kernel)


# ---- group_gemm.py --------------------------------------------------------------------------------------------
def moe_grouped_gemm_2weights(x_sorted, w_gate, w_up, routing, split: str = "N"):
    """(group_gemm.py:318/:402) two weights sharing the A operand: run as one grouped GEMM on the stacked weight."""
    w = torch.cat([w_gate, w_up], dim=1)
    return M.moe_grouped_gemm(x_sorted, w, routing)


transposed_moe_grouped_gemm = M.transposed_moe_grouped_gemm      # one launch, segmented-K batch mode (ops/moe.py)


def calc_gather_scatter_index_triton(topk_ids: torch.Tensor, num_experts: int, block_m: int = 128):
    """(moe_utils.py:308) -> (gather_index = sorted flat ids, expert tile ids, padded offsets)."""
    r = M.moe_align_sort(topk_ids, num_experts, block_m)
    return r.sorted_ids, r.tile_expert, r.expert_offsets


histogram_by_expert_triton = M.histogram_by_expert
reduce_topk_tma = reduce_topk_non_tma = M.reduce_topk


# ---- all_to_all_single_gemm.py / Ulysses GEMM fusions ----------------------------------------------------------
def create_all_to_all_single_gemm_context(max_rows_per_peer: int, K: int, dtype: torch.dtype, N: int = 0):
    """Workspace + per-source arrival flags of the fused AllToAll+GEMM (the all-gather GEMM context: same protocol)."""
    from .ag_gemm import create_ag_gemm_context
    heap = U.get_heap()
    return create_ag_gemm_context(max_rows_per_peer * heap.world, N, K, dtype, heap.rank, heap.world)


def all_to_all_single_gemm(ctx, x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(all_to_all_single_gemm.py:74-188) ``x``: [W * rows, K], row block d goes to rank d; returns
    ``concat_s(block received from s) @ w.T``.  ONE kernel: comm CTAs push block d into rank d's workspace and raise
    per-(source, slice) flags, the tcgen05 GEMM tiles of the same launch wait on the flags of the source they read."""
    from .ag_gemm import ag_gemm
    return ag_gemm(x, w.t(), ctx, out=out, all_to_all=True)


def gemm_only(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    return _lin(x, w)


class SpUlysessQKVGemmAll2AllKernel:
    """(sp_ulysess_qkv_gemm_all2all.py:545-963) QKV projection of the local sequence shard, then head<->sequence
    all-to-all: ``[S/W, H] -> q,k,v [S, heads/W, D]``."""

    def __init__(self, max_local_seq: int, num_q_heads: int, num_kv_heads: int, head_dim: int, dtype, rank: int, world_size: int):
        from ..parallel.sp import UlyssesSPAllToAllLayer
        self.Hq, self.Hkv, self.D, self.W = num_q_heads, num_kv_heads, head_dim, world_size
        from ..parallel.sp import UlyssesQKVPackAllToAll
        self.a2a_q = UlyssesSPAllToAllLayer(max_local_seq, num_q_heads, head_dim, dtype, rank, world_size)
        self.a2a_kv = UlyssesSPAllToAllLayer(max_local_seq, num_kv_heads, head_dim, dtype, rank, world_size)
        self.pack = UlyssesQKVPackAllToAll(max_local_seq, num_q_heads, num_kv_heads, head_dim, dtype, rank, world_size)

    def forward(self, x: torch.Tensor, wqkv: torch.Tensor):
        """QKV GEMM on the tcgen05 kernel, then q, k, v in ONE packed all-to-all (the fully fused alternative -- the all-to-all in
        the GEMM epilogue -- is :func:`triton_dist.ops.gemm_a2a.gemm_all_to_all`, see ``ulysses_sp_infer_gemm_a2a_op``)."""
        qkv = _lin(x, wqkv)
        S = x.shape[0]
        q, k, v = qkv.split([self.Hq * self.D, self.Hkv * self.D, self.Hkv * self.D], dim=-1)
        return self.pack(q.reshape(S, self.Hq, self.D), k.reshape(S, self.Hkv, self.D), v.reshape(S, self.Hkv, self.D))

    pre_attn_a2a = qkv_pack_a2a = forward

    def finalize(self):
        self.a2a_q.finalize(); self.a2a_kv.finalize(); self.pack.finalize()


class SpUlysessOAll2AllGemmKernel:
    """(sp_ulysess_o_all2all_gemm.py) attention output ``[S, heads/W, D]`` -> all-to-all -> ``[S/W, heads * D] @ wo.T``."""

    def __init__(self, max_local_seq: int, num_q_heads: int, head_dim: int, dtype, rank: int, world_size: int):
        from ..parallel.sp import UlyssesSPAllToAllLayer
        self.a2a = UlyssesSPAllToAllLayer(max_local_seq, num_q_heads, head_dim, dtype, rank, world_size)

    def forward(self, attn_out: torch.Tensor, wo: torch.Tensor) -> torch.Tensor:
        o = self.a2a.post_attn_a2a(attn_out)
        return _lin(o.reshape(o.shape[0], -1).contiguous(), wo)

    post_attn_a2a = forward

    def finalize(self):
        self.a2a.finalize()


UlyssesSpInferPreAttnContext = SpUlysessQKVGemmAll2AllKernel


def ulysses_sp_infer_gemm_a2a_op(ctx, x, wqkv, input_scale=None, weight_scale=None):
    """(ulysses_sp_infer_gemm_a2a.py:455) ``ctx``: a :class:`SpUlysessQKVGemmAll2AllKernel` (GEMM then all-to-all) or a
    :class:`triton_dist.ops.gemm_a2a.GemmA2AContext` (all-to-all fused into the GEMM epilogue; ``wqkv`` rows grouped by
    destination rank).  int8 / float8_e4m3fn ``x`` and ``wqkv`` with ``input_scale`` (per row) / ``weight_scale`` (per output
    channel) run the quantised flavour (dequantisation in the tcgen05 epilogue, bf16 on the wire)."""
    from .gemm_a2a import GemmA2AContext, gemm_all_to_all
    if isinstance(ctx, GemmA2AContext):
        return gemm_all_to_all(ctx, x, wqkv, scale_a=input_scale, scale_b=weight_scale)
    if input_scale is not None or weight_scale is not None:
        raise NotImplementedError("quantised operands need a GemmA2AContext (create_gemm_a2a_context)")
    return ctx.forward(x, wqkv)


# ---- ep_a2a.py (normal mode) ----------------------------------------------------------------------------------
def ep_dispatch_token_inplace(layer, x, topk_idx, topk_weights=None):
    """(ep_a2a.py:881) throughput-mode dispatch.  ``layer``: :class:`triton_dist.parallel.ep.EPNormalAll2AllLayer` (token saving,
    index-list receive side) or the low-latency layer (same NVLink push protocol with bf16 payloads)."""
    from ..parallel.ep import EPNormalAll2AllLayer
    if isinstance(layer, EPNormalAll2AllLayer):
        return layer.dispatch(x, topk_idx, topk_weights)
    return layer.dispatch(x, None, topk_idx)


def ep_combine_token_inplace(layer, expert_out, topk_idx, topk_weights, meta):
    from ..parallel.ep import EPNormalAll2AllLayer
    if isinstance(layer, EPNormalAll2AllLayer):
        return layer.combine(expert_out, meta, topk_idx)
    return layer.combine(expert_out, topk_idx, topk_weights, meta)


def get_ag_splits_and_recv_offset_for_dispatch(topk_idx: torch.Tensor, num_experts: int):
    """(ep_a2a.py:765) per-expert token counts on this rank (the receive offsets are produced inside dispatch)."""
    return M.histogram_by_expert(topk_idx, num_experts)


# ---- ulysses_sp_dispatch.py -----------------------------------------------------------------------------------
def create_ulysses_sp_pre_attn_comm_context(max_local_seq: int, num_q_heads: int, num_kv_heads: int, head_dim: int, dtype,
                                            rank: Optional[int] = None, world_size: Optional[int] = None):
    """(ulysses_sp_dispatch.py:546) one context holding the q and kv all-to-all workspaces."""
    heap = U.get_heap()
    r = heap.rank if rank is None else rank
    w = heap.world if world_size is None else world_size
    return SpUlysessQKVGemmAll2AllKernel(max_local_seq, num_q_heads, num_kv_heads, head_dim, dtype, r, w)


def pre_attn_qkv_pack_a2a_op(ctx: SpUlysessQKVGemmAll2AllKernel, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
    """(ulysses_sp_dispatch.py:606) seq-sharded ``[S/W, H, D]`` q/k/v -> head-sharded ``[S, H/W, D]``: ONE packed all-to-all."""
    return ctx.pack(q, k, v)


def qkv_bsnd_to_bnsd(x: torch.Tensor) -> torch.Tensor:
    """(ulysses_sp_dispatch.py:411) ``[B, S, N, D] -> [B, N, S, D]`` contiguous."""
    out = torch.empty((x.shape[0], x.shape[2], x.shape[1], x.shape[3]), dtype=x.dtype, device=x.device)
    out.copy_(x.transpose(1, 2))
    return out


# ---- all_to_all_vdev_2d_offset.py -------------------------------------------------------------------------------
def all_to_all_vdev_2d_offset(ctx: AllToAllContext, send: torch.Tensor, in_splits: torch.Tensor, in_offsets: torch.Tensor,
                              experts_per_rank: int):
    """(all_to_all_vdev_2d_offset.py) variable all-to-all where the rows for (dst rank d, local expert e) start at
    ``in_offsets[d * epr + e]`` (the rows are not packed).  The send buffer is compacted with ONE device-side gather whose index
    is built without reading the splits on the host (positions beyond the total are dummies the kernel never sends), then the
    device-split all-to-all kernel runs; returns ``(recv, recv_splits[W, epr])``."""
    from .all_to_all import all_to_all_vdev_2d
    sp = in_splits.to(torch.int64)
    cum = torch.cumsum(sp, 0)
    R = send.shape[0]                                   # static upper bound of the packed length: no .item()
    pos = torch.arange(R, device=send.device)
    seg = torch.searchsorted(cum, pos, right=True).clamp(max=sp.numel() - 1)
    rows = torch.where(pos < cum[-1], in_offsets.to(torch.int64)[seg] + pos - (cum - sp)[seg], torch.zeros_like(pos))
    return all_to_all_vdev_2d(ctx, send.index_select(0, rows.clamp(0, R - 1)).contiguous(), in_splits)


# ---- sp_ag_attention_inter_node.py --------------------------------------------------------------------------------
def fused_sp_ag_attn_inter_node(ctx, q_shard, k_shard, v_shard, **kw):
    """Multi-node variant of the context-parallel prefill.  This framework targets one NVSwitch domain (<= 72 GPUs on
    NVL72 are one "node" for the symmetric heap), so it is the intra-node path (sp_ag_attention_inter_node.py:116-190)."""
    from ..parallel.sp import fused_sp_ag_attn_intra_node
    return fused_sp_ag_attn_intra_node(ctx, q_shard, k_shard, v_shard, **kw)


# ---- ep_all2all_fused.py (Mega-EP entry points) -------------------------------------------------------------------
def mega_kernel_dispatch_token_moe_grouped_gemm(op, x: torch.Tensor, topk_idx: torch.Tensor, topk_w: torch.Tensor, w_gate_up=None):
    """(ep_all2all_fused.py:839) dispatch fused with the first grouped GEMM (+ SwiGLU): returns the activations in the expert-sorted
    layout and the handle the second half needs.  ``op``: a :class:`triton_dist.parallel.ep.EpAll2AllFusedOp` (ONE kernel,
    csrc/gemm_sm100.cuh mode kEPD; pass ``w_gate_up``) or an :class:`~triton_dist.parallel.ep.EP_MoE` layer (its own weights)."""
    if hasattr(op, "mega_dispatch_group_gemm"):
        return op.mega_dispatch_group_gemm(x, topk_idx, topk_w, w_gate_up)
    return op.dispatch_group_gemm(x, topk_idx, topk_w)


def mega_kernel_moe_grouped_gemm_combine_token(op, act: torch.Tensor, handle, w_down=None):
    """(ep_all2all_fused.py:1020) second grouped GEMM fused with the combine transfer (mode kEPC) + weighted top-k reduce on the owner."""
    if hasattr(op, "mega_group_gemm_combine"):
        return op.mega_group_gemm_combine(act, handle, w_down)
    return op.group_gemm_combine(act, handle)


# ---- low_latency_allgather.py: one entry point per method name (reference :819-960) -------------------------------
def _fast_ag(mode):
    def f(ctx, symm_buffer: torch.Tensor, output: Optional[torch.Tensor] = None):
        from . import comm
        return comm.fast_allgather(symm_buffer, ctx, mode=mode, output=output)
    f.__name__ = f"fast_allgather_{mode}"
    f.__doc__ = f"``fast_allgather(shard, ctx, mode={mode!r})`` (kernel table: ops/comm.py ``_AG_MODES``)."
    return f


fast_allgather_pull = _fast_ag("pull")
fast_allgather_push_2d = _fast_ag("push_2d")
fast_allgather_push_3d = _fast_ag("push_3d")
fast_allgather_push_2d_ll = _fast_ag("push_2d_ll")
fast_allgather_push_2d_ll_multimem = _fast_ag("push_2d_ll_multimem")
fast_allgather_push_numa_2d = _fast_ag("push_numa_2d")
fast_allgather_push_numa_2d_ll = _fast_ag("push_numa_2d_ll")
fast_allgather_push_multimem = _fast_ag("push_multimem")


def fast_allgather_push_numa_2d_ll_multinode(*_a, **_k):
    raise NotImplementedError("multi-node all-gather is out of scope (single NVSwitch domain); use fast_allgather_push_numa_2d_ll")


# ---- allgather.py copy-engine producers by name -------------------------------------------------------------------
def cp_engine_producer_all_gather_full_mesh_push(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, stream=None, **kw):
    from .allgather import cp_engine_producer_all_gather_intra_node
    return cp_engine_producer_all_gather_intra_node(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, **kw)


def cp_engine_producer_all_gather_full_mesh_pull(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, stream=None,
                                                 signal_value: int = 1, **_):
    """Pull flavour: publish my shard in my own buffer, raise my flag on every peer, then copy every peer's shard once its flag
    arrived (the push producer of ops/allgather.py writes into the peers instead)."""
    from .. import language as dl
    M = local_tensor.shape[0]
    remote_tensor_buffers[rank][rank * M:(rank + 1) * M].copy_(local_tensor)
    for q in range(1, num_ranks):
        dl.notify(barrier_buffers[rank][rank:rank + 1], (rank + q) % num_ranks, signal=signal_value, sig_op="set")
    barrier_buffers[rank][rank:rank + 1].fill_(signal_value)
    for q in range(1, num_ranks):
        src = (rank + q) % num_ranks
        dl.wait(barrier_buffers[rank][src:src + 1], 1, wait_value=signal_value)
        remote_tensor_buffers[rank][src * M:(src + 1) * M].copy_(remote_tensor_buffers[src][src * M:(src + 1) * M])


def cp_engine_producer_all_gather_ring_push_numa_2d(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, **kw):
    """The NUMA-staged ring exists for PCIe / multi-socket topologies; inside one NVSwitch domain it is the 2-D ring."""
    from .allgather import cp_engine_producer_all_gather_ring_push_2d
    return cp_engine_producer_all_gather_ring_push_2d(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, **kw)


# ---- context classes under the reference's names ------------------------------------------------------------------
def _ctx_aliases():
    from . import all_to_all as _a2a, comm as _comm, ep_a2a as _ep, gemm_a2a as _ga, gemm_ar as _gar
    from ..parallel import sp as _sp
    return dict(AllToAllSingle2DContext=_a2a.AllToAllContext, AllToAllSingleGemmContext=_ga.GemmA2AContext,
                MoEAllGatherGroupGEMMTensorParallelContext=M.MoEAllGatherGroupGEMMContext, MoEReduceARContext=M.MoEReduceRSContext,
                ReduceScatter2DContext=_comm.AllReduceContext, LLGemmARContext=_gar.GemmARContext, EPContext=_ep.EPLowLatencyContext,
                LowlatencyDispatchContext=_ep.EPLowLatencyContext, LowlatencyCombineContext=_ep.EPLowLatencyContext,
                SPAllGatherAttentionContextInterNode=_sp.SPAllGatherAttentionContextIntraNode,
                UlyssesSPPreAttnCommContext=SpUlysessQKVGemmAll2AllKernel)


def __getattr__(name):                      # resolved lazily: the context classes live in modules that import this one
    table = _ctx_aliases()
    if name in table:
        return table[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def create_sp_ag_attention_context_inter_node(*a, **k):
    from ..parallel.sp import create_sp_ag_attention_context_intra_node
    return create_sp_ag_attention_context_intra_node(*a, **k)


def gemm_rs_op(A, B, ctx, *a, **k):
    """(gemm_reduce_scatter.py ``gemm_rs_op``) the functional spelling of :func:`triton_dist.ops.gemm_rs.gemm_rs`."""
    from .gemm_rs import gemm_rs
    return gemm_rs(A, B, ctx, *a, **k)


# ---- moe_utils.py torch goldens -----------------------------------------------------------------------------------
def histogram_by_expert_torch(topk_ids: torch.Tensor, num_experts: int) -> torch.Tensor:
    return torch.bincount(topk_ids.reshape(-1).long().clamp_min(0), minlength=num_experts)[:num_experts].to(torch.int32)


def calc_scatter_index_torch(topk_ids: torch.Tensor, num_experts: int) -> torch.Tensor:
    """scatter_index[t, k] = row of pair (t, k) in the expert-sorted order (stable)."""
    flat = topk_ids.reshape(-1).long()
    order = torch.sort(flat, stable=True).indices
    scatter = torch.empty_like(order)
    scatter[order] = torch.arange(flat.numel(), device=flat.device)
    return scatter.view_as(topk_ids).to(torch.int32)


def calc_gather_index_from_scatter_index(scatter_index: torch.Tensor) -> torch.Tensor:
    flat = scatter_index.reshape(-1).long()
    gather = torch.empty_like(flat)
    gather[flat] = torch.arange(flat.numel(), device=flat.device)
    return gather.to(torch.int32)


def calc_gather_index_torch(topk_ids: torch.Tensor, num_experts: int) -> torch.Tensor:
    """gather_index[i] = flat (token * topk + k) index of the i-th row of the expert-sorted order."""
    return calc_gather_index_from_scatter_index(calc_scatter_index_torch(topk_ids, num_experts))


# ---- names of the reference that map onto one implementation here ---------------------------------------------------------------
from .ep_metadata import (get_ag_splits_and_recv_offset_for_dispatch_intra_node, get_dispatch_send_reqs,  # noqa: E402,F401
                          recv_offsets_from_splits)
from .tile_swizzle import (threadblock_swizzle_allgather_gemm, threadblock_swizzle_allgather_gemm_kernel,  # noqa: E402,F401
                           threadblock_swizzle_gemm_reduce_scatter, threadblock_swizzle_gemm_reduce_scatter_kernel)

calc_gather_scatter_index_v2_triton = calc_gather_scatter_index_triton      # (moe_utils.py:341) one implementation: the align/sort kernel


def run_moe_reduce_rs_triton_non_overlap(x, w, chosen_experts, expert_weight, ctx=None, group=None, **_hints):
    """(moe_reduce_rs.py:825) the NON-overlapped baseline the fused op is compared against: grouped GEMM, weighted top-k reduce, then
    one NCCL ``reduce_scatter_tensor`` -- three steps back to back, no chunking, no symmetric buffers."""
    W, me = U.world_size(), U.rank()
    return M.moe_reduce_rs_torch(x, w, chosen_experts, expert_weight, group or U.get_triton_dist_world(), W, me)


def run_moe_reduce_ar_triton_non_overlap(x, w, chosen_experts, expert_weight, ctx=None, group=None, **_hints):
    """(moe_reduce_ar.py) the non-overlapped baseline of ``run_moe_reduce_ar``: grouped GEMM, weighted top-k reduce, one all-reduce."""
    return M.moe_reduce_ar_torch(x, w, chosen_experts, expert_weight, group or U.get_triton_dist_world(), U.world_size())


def create_context(max_num_token: int, token_len_elem: int, num_expert_per_rank: int = 1, dtype: torch.dtype = torch.bfloat16, **_hints):
    """(all_to_all_vdev_2d_offset.py:528) context of the variable-size 2-D all-to-all: rows of ``token_len_elem`` elements, at most
    ``max_num_token`` rows per peer."""
    from .all_to_all import create_all_to_all_context
    W = U.world_size()
    return create_all_to_all_context(max_num_token, token_len_elem, U.rank(), num_expert_per_rank * W, W, num_expert_per_rank, dtype)


def all_to_all_v_offset_op(ctx, rank_in_row: bool = True, input: torch.Tensor = None, output: torch.Tensor = None,
                           in_splits: torch.Tensor = None, in_offset: torch.Tensor = None, has_input_offset: bool = False, **_hints):
    """(all_to_all_vdev_2d_offset.py:637) the common entry of ``all_to_all_vdev_2d`` (packed rows) and ``..._offset`` (rows of every
    (destination, expert) group start at ``in_offset``).  Returns ``(received rows, recv_splits)``; ``output`` is filled when given."""
    from .all_to_all import all_to_all_vdev_2d
    epr = ctx.experts_per_rank if hasattr(ctx, "experts_per_rank") else max(1, in_splits.numel() // U.world_size())
    if has_input_offset or in_offset is not None:
        recv, splits = all_to_all_vdev_2d_offset(ctx, input, in_splits.reshape(-1), in_offset.reshape(-1), epr)
    else:
        recv, splits = all_to_all_vdev_2d(ctx, input, in_splits.reshape(-1))
    if output is not None:
        output[: recv.shape[0]].copy_(recv)
    return recv, splits


all_to_all_v_offset_op_v2 = all_to_all_v_offset_op

# flash_decode.py *_aot: every kernel of this framework is compiled ahead of time by nvcc, the AOT entry points are the same functions


def pre_attn_a2a_comm_only(ctx, qkv_local: torch.Tensor):
    """(ulysses_sp_infer_gemm_a2a.py:394) the all-to-all half of the inference Ulysses op without the GEMM: ``qkv_local``
    [S / W, (Hq + 2 Hkv) * D] seq-sharded projections -> head-sharded (q, k, v) [S, H / W, D]."""
    S_loc = qkv_local.shape[0]
    q, k, v = qkv_local.view(S_loc, -1, ctx.D).split([ctx.Hq, ctx.Hkv, ctx.Hkv], dim=1)
    return ctx.pack(q.contiguous(), k.contiguous(), v.contiguous())


# ---- common_ops.py device barriers: one symmetric-heap barrier here (csrc/td/primitives.cuh ``barrier_all_block``: flag flip on a
# monotone epoch, the non-atomic protocol of the reference :172-224; the CAS variant :154-168 exists there for hardware without native
# P2P atomics ordering -- NVLink 5 needs neither) --------------------------------------------------------------------------------------
def __getattr__(name, _prev=globals().get("__getattr__")):
    if name in ("barrier_all_intra_node_atomic_cas_block", "barrier_all_intra_node_non_atomic", "barrier_all_intra_node_non_atomic_block"):
        from ..lk import ll
        return ll.barrier_all_block
    if _prev is not None:
        return _prev(name)
    raise AttributeError(name)

