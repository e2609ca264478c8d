"""MoE building blocks: routing sort/align, grouped GEMM on tcgen05, top-k reduce, and the two TP-MoE fused ops.

Reference:
  * ``calc_gather_scatter_index`` / ``reduce_topk`` -- kernels/nvidia/moe_utils.py:145-508
  * ``moe_grouped_gemm`` -- kernels/nvidia/group_gemm.py:251
  * ``create_ag_group_gemm_context`` / ``ag_group_gemm`` -- kernels/nvidia/allgather_group_gemm.py:338-609
  * ``create_moe_rs_context`` / ``run_moe_reduce_rs`` -- kernels/nvidia/moe_reduce_rs.py:88, 872-961
  * ``create_moe_ar_context`` / ``run_moe_reduce_ar`` -- kernels/nvidia/moe_reduce_ar.py:645

B200 design: tokens are sorted by expert once on the device (one CTA counting sort, segments padded to the tile
height so an m-tile never mixes experts), gathered into expert-contiguous rows, and the *same* tcgen05 GEMM kernel
runs with a per-m-tile expert id that offsets the B (weight) TMA coordinate -- no separate grouped-GEMM kernel and
no host sync: the padded row count stays on the device and unused tiles are skipped by their ``-1`` expert id.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .. import _C
from .. import utils as U
from . import comm
from .gemm import GemmConfig, fill_common

c_void_p, c_int, c_ll = C.c_void_p, C.c_int, C.c_longlong
_C.register("td_moe_align_sort", c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p])
_C.register("td_gather_rows", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_void_p])
_C.register("td_scatter_rows", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p])
_C.register("td_topk_reduce", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p])
_C.register("td_bincount", c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p])
_C.register("td_transpose_gather", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p])


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------------------
# routing
# ------------------------------------------------------------------------------------------------------------
@dataclass
class SortedRouting:
    sorted_ids: torch.Tensor      # [capacity] flat (token * topk + k) index, or pad_id
    tile_expert: torch.Tensor     # [capacity / block_m]
    expert_offsets: torch.Tensor  # [E + 1]
    total_padded: torch.Tensor    # [1] (device)
    capacity: int
    block_m: int
    pad_id: int


def padded_capacity(n_pairs: int, E: int, block_m: int) -> int:
    """Upper bound of the padded row count: every expert wastes < block_m rows."""
    return (n_pairs + E * (block_m - 1) + block_m - 1) // block_m * block_m


def moe_align_sort(topk_ids: torch.Tensor, num_experts: int, block_m: int = 128, tokens_per_rank: int = 0,
                   rank: int = 0, world: int = 1) -> SortedRouting:
    """Sort the flat (token, k) pairs by expert, pad each expert segment to ``block_m`` rows.  With
    ``tokens_per_rank`` the pairs of each expert are additionally ordered by all-gather arrival stage of their source
    rank (the reference's threadblock_swizzle_ag_moe: tiles whose tokens arrive first run first)."""
    ids = topk_ids.reshape(-1).to(torch.int32).contiguous()
    n = ids.numel()
    topk = topk_ids.shape[-1] if topk_ids.dim() > 1 else 1
    cap = padded_capacity(n, num_experts, block_m)
    dev = ids.device
    pad_id = n
    if not ids.is_cuda:
        stage = ((torch.arange(n) // topk) // tokens_per_rank - rank) % world if tokens_per_rank > 0 else torch.zeros(n, dtype=torch.long)
        valid = (ids >= 0) & (ids < num_experts)                         # unrouted slots (id < 0) are dropped, like the CUDA kernel
        key = torch.where(valid, ids.long(), torch.full_like(ids.long(), num_experts)) * (world + 1) * (n + 1) + stage * (n + 1) + torch.arange(n)
        order = torch.argsort(key)
        counts = torch.bincount(ids.long()[valid], minlength=num_experts)
        padded = (counts + block_m - 1) // block_m * block_m
        offs = torch.zeros(num_experts + 1, dtype=torch.int64)
        offs[1:] = torch.cumsum(padded, 0)
        sorted_ids = torch.full((cap,), pad_id, dtype=torch.int32)
        tile_expert = torch.full((cap // block_m,), -1, dtype=torch.int32)
        src = 0
        for e in range(num_experts):
            c = int(counts[e])
            sorted_ids[int(offs[e]):int(offs[e]) + c] = order[src:src + c].to(torch.int32)
            tile_expert[int(offs[e]) // block_m:int(offs[e + 1]) // block_m] = e
            src += c
        return SortedRouting(sorted_ids, tile_expert, offs.to(torch.int32), offs[-1:].to(torch.int32), cap, block_m, pad_id)
    sorted_ids = torch.empty(cap, dtype=torch.int32, device=dev)
    tile_expert = torch.empty(cap // block_m, dtype=torch.int32, device=dev)
    offs = torch.empty(num_experts + 1, dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    buckets = num_experts * (world if tokens_per_rank > 0 else 1)
    ws = torch.empty(buckets * ((n + 1023) // 1024 + 1) + buckets, dtype=torch.int32, device=dev)     # histogram / scan scratch
    _C.check(_C.cuda_lib().td_moe_align_sort(ids.data_ptr(), n, num_experts, block_m, cap, pad_id, sorted_ids.data_ptr(),
                                             tile_expert.data_ptr(), offs.data_ptr(), total.data_ptr(), topk, tokens_per_rank,
                                             rank, world, ws.data_ptr(), ws.numel() * 4, _s()), "td_moe_align_sort")
    return SortedRouting(sorted_ids, tile_expert, offs, total, cap, block_m, pad_id)


def gather_rows(src: torch.Tensor, routing: SortedRouting, div: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[i] = src[sorted_ids[i] // div]`` (zeros for pad rows)."""
    H = src.shape[-1]
    out = torch.empty((routing.capacity, H), dtype=src.dtype, device=src.device) if out is None else out
    if not src.is_cuda:
        ids = routing.sorted_ids.long()
        valid = ids != routing.pad_id
        out.zero_()
        out[valid] = src[ids[valid] // div]
        return out
    _C.check(_C.cuda_lib().td_gather_rows(out.data_ptr(), src.contiguous().data_ptr(), routing.sorted_ids.data_ptr(), None,
                                          routing.capacity, H * src.element_size(), div, routing.pad_id, _s()), "td_gather_rows")
    return out


def scatter_rows(src_sorted: torch.Tensor, routing: SortedRouting, n_rows: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[sorted_ids[i]] = src_sorted[i]`` -- back to flat (token, k) order."""
    H = src_sorted.shape[-1]
    out = torch.empty((n_rows, H), dtype=src_sorted.dtype, device=src_sorted.device) if out is None else out
    if not src_sorted.is_cuda:
        ids = routing.sorted_ids.long()
        valid = ids != routing.pad_id
        out[ids[valid]] = src_sorted[valid]
        return out
    _C.check(_C.cuda_lib().td_scatter_rows(out.data_ptr(), src_sorted.data_ptr(), routing.sorted_ids.data_ptr(), None,
                                           routing.capacity, H * src_sorted.element_size(), routing.pad_id, _s()), "td_scatter_rows")
    return out


def reduce_topk(y: torch.Tensor, weights: Optional[torch.Tensor], topk: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[t] = sum_j weights[t, j] * y[t * topk + j]`` with fp32 accumulation (moe_utils.py:reduce_topk_*)."""
    T = y.shape[0] // topk
    H = y.shape[1]
    out = torch.empty((T, H), dtype=y.dtype, device=y.device) if out is None else out
    if not y.is_cuda or H % 8 or y.dtype not in (torch.bfloat16, torch.float16):
        w = weights.float().reshape(T, topk, 1) if weights is not None else 1.0
        out.copy_((y.float().view(T, topk, H) * w).sum(1).to(y.dtype))
        return out
    wf = weights.float().contiguous() if weights is not None else None
    _C.check(_C.cuda_lib().td_topk_reduce(out.data_ptr(), y.contiguous().data_ptr(), wf.data_ptr() if wf is not None else None,
                                          T, topk, H, int(y.dtype == torch.bfloat16), _s()), "td_topk_reduce")
    return out


def histogram_by_expert(topk_ids: torch.Tensor, num_experts: int) -> torch.Tensor:
    ids = topk_ids.reshape(-1).to(torch.int32).contiguous()
    if not ids.is_cuda:
        return torch.bincount(ids.long(), minlength=num_experts).to(torch.int32)
    out = torch.empty(num_experts, dtype=torch.int32, device=ids.device)
    _C.check(_C.cuda_lib().td_bincount(ids.data_ptr(), ids.numel(), out.data_ptr(), num_experts, _s()), "td_bincount")
    return out


bincount = histogram_by_expert


# ------------------------------------------------------------------------------------------------------------
# grouped GEMM
# ------------------------------------------------------------------------------------------------------------
def moe_grouped_gemm_fused(src: torch.Tensor, w: torch.Tensor, routing: SortedRouting, div: int, n_out_rows: int,
                           out: Optional[torch.Tensor] = None, config: Optional[GemmConfig] = None,
                           n_slice: Optional[Tuple[int, int]] = None, gather_idx: Optional[torch.Tensor] = None,
                           gather: bool = True, scatter: bool = True) -> torch.Tensor:
    """``out[id] = src[id // div] @ w[expert(id)].T`` for every routed pair id = token * topk + k, in ONE kernel: the A rows
    of each tile are fetched with TMA ``tile::gather4`` straight from ``src`` (no gather_rows pass) and the epilogue writes
    each row to its final place (no scatter_rows pass).  The sm_100a counterpart of the reference's gather/scatter
    grouped GEMM (allgather_group_gemm.py:536-609, which cannot use TMA for the gathered operand).

    ``gather_idx`` (int32 [capacity], -1 = zero row) replaces ``sorted_ids // div`` as the source row of every sorted position
    (rows that are addressed through a second level of indirection, e.g. EP receive buffers); ``gather=False``: ``src`` already is
    the sorted, padded layout; ``scatter=False``: the output stays in the sorted layout ``[capacity, N]``."""
    K = src.shape[1]
    E, N_full, Kw = w.shape
    assert Kw == K and src.is_cuda and src.stride(1) == 1 and w.is_contiguous()
    n0, N = n_slice if n_slice is not None else (0, N_full)          # columns [n0, n0 + N) of every expert's weight
    assert not (gather_idx is not None and scatter), "an explicit gather index and the scatter epilogue use different pad ids"
    if not scatter:
        n_out_rows = routing.capacity
    out = torch.empty((n_out_rows, N), dtype=src.dtype, device=src.device) if out is None else out
    cg = routing.block_m // 128
    cfg = config or GemmConfig(bn=256 if N >= 256 else 128, cta_group=cg, group_m=1, use_tma_store=False)
    assert routing.block_m in (128, 256) and cfg.cta_group == cg, "routing block_m must equal the GEMM tile height (128 * cta_group)"
    assert cg == 1 or not gather, "the TMA gather4 producer runs with cta_group 1; CTA pairs take pre-sorted rows (gather=False)"
    args = _C.GemmArgs()
    args.mode = 0
    w2 = w.reshape(E * N_full, K)
    fill_common(args, src.shape[0], src.data_ptr(), src.stride(0), w2, out.data_ptr(), n_out_rows, out.stride(0),
                routing.capacity, N, K, GemmConfig(cfg.bn, cg, 1, False, cfg.num_sms, 0), src.dtype == torch.bfloat16)
    args.B = w2.data_ptr() + n0 * K * w.element_size()
    args.expert_stride_rows = N_full
    args.tile_expert, args.num_experts = routing.tile_expert.data_ptr(), E
    if gather:
        if gather_idx is not None:
            assert gather_idx.dtype == torch.int32 and gather_idx.numel() == routing.capacity and gather_idx.is_contiguous()
            args.a_gather, args.a_gather_div, args.a_gather_pad = gather_idx.data_ptr(), 1, -1
        else:
            args.a_gather, args.a_gather_div, args.a_gather_pad = routing.sorted_ids.data_ptr(), div, routing.pad_id
        args.a_src_rows = src.shape[0]
    else:
        assert src.shape[0] == routing.capacity
        args.a_gather_pad = routing.pad_id
    if scatter:
        args.c_scatter = routing.sorted_ids.data_ptr()
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), _s()), "td_gemm_launch(grouped, gather4)")
    return out


_PAIRS_FOR_CTA_PAIRS = 4096     # from this many routed pairs on, the grouped GEMM is compute bound: pre-sort rows, use 2-CTA tiles


def moe_grouped_gemm_presorted(src: torch.Tensor, w: torch.Tensor, topk_ids: torch.Tensor, num_experts: int, div: int,
                               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Large routed batches: one memory-bound ``gather_rows`` pass puts the rows in expert order (padded to 256), then the
    full-rate 2-CTA 256x256 tcgen05 tiles run on plain tiled TMA and the epilogue scatters straight to (token, k) order.
    Measured on B200 (Mixtral down projection, 16 K pairs): the tile::gather4 producer issues 32 TMA instructions per
    k-block and runs at ~0.45 PFLOP/s, this path at the plain-GEMM rate; gather4 stays the choice for small (decode) batches."""
    r = moe_align_sort(topk_ids, num_experts, 256)
    xs = gather_rows(src, r, div=div)
    return moe_grouped_gemm_fused(xs, w, r, div, topk_ids.numel(), out=out, gather=False, scatter=True)


def transpose_gather(src: torch.Tensor, ids: Optional[torch.Tensor], n_out: int) -> torch.Tensor:
    """``out[c, i] = src[ids[i], c]`` (zero column where ``ids[i] < 0``; identity if ``ids`` is None): token-major matrix ->
    reduction-major operand of a weight-gradient GEMM, one memory-bound pass."""
    rows, cols = src.shape
    out = torch.empty((cols, n_out), dtype=src.dtype, device=src.device)
    if not src.is_cuda or cols % 8 or n_out % 8 or src.stride(0) % 8 or src.stride(1) != 1:
        idx = torch.arange(n_out, device=src.device) if ids is None else ids.long()
        ok = (idx >= 0) & (idx < rows)
        g = torch.zeros((n_out, cols), dtype=src.dtype, device=src.device)
        g[ok] = src[idx[ok]]
        out.copy_(g.t())
        return out
    _C.check(_C.cuda_lib().td_transpose_gather(out.data_ptr(), src.data_ptr(), ids.data_ptr() if ids is not None else None, n_out, cols,
                                               src.stride(0), _s()), "td_transpose_gather")
    return out


def transposed_moe_grouped_gemm(grad_output: torch.Tensor, original_input: torch.Tensor, split_size: torch.Tensor,
                                split_size_cum_per_expert: Optional[torch.Tensor] = None, grad_weight: Optional[torch.Tensor] = None,
                                **ref_hints) -> torch.Tensor:
    """Weight gradient of a grouped GEMM (reference kernels/nvidia/group_gemm.py:503-727, :988):
    ``grad_weight[g] = grad_output[rows of g].T @ original_input[rows of g]`` for expert-contiguous rows with ``split_size[g]`` rows
    each.  Both operands are token-major, i.e. MN-major for this product: one ``transpose_gather`` pass per operand lays them out
    reduction-major with every expert's segment padded to 64 tokens (zero columns), then ONE launch of the tcgen05 GEMM in its
    segmented-K batch mode computes all experts (batch g reduces over the k-blocks of its own segment; csrc/gemm_sm100.cuh
    ``segk_off``).  No per-expert Python loop, no host sync."""
    U.accept_ref_hints("transposed_moe_grouped_gemm", ref_hints, ("BLOCK_SIZE_M", "BLOCK_SIZE_N", "BLOCK_SIZE_K", "GROUP_SIZE_M", "num_warps",
                                                                  "num_stages", "persistent", "sm_margin"))
    M_, N = grad_output.shape
    M2, K = original_input.shape
    G = split_size.numel()
    assert M_ == M2 and grad_output.dtype == original_input.dtype
    dev = grad_output.device
    split = split_size.to(torch.int64)
    cum = torch.cumsum(split, 0) - split if split_size_cum_per_expert is None else (split_size_cum_per_expert.to(torch.int64) - split)
    if grad_weight is None:
        grad_weight = torch.zeros((G, N, K), dtype=grad_output.dtype, device=dev)
    else:
        assert grad_weight.shape == (G, N, K) and grad_weight.is_contiguous()
        grad_weight.zero_()                       # batches with an empty segment are skipped by the kernel
    if not grad_output.is_cuda or N % 8 or K % 8:
        for g in range(G):
            s, n = int(cum[g]), int(split[g])
            if n:
                grad_weight[g] = (grad_output[s:s + n].float().t() @ original_input[s:s + n].float()).to(grad_weight.dtype)
        return grad_weight
    pad = (split + 63) // 64 * 64
    pend = torch.cumsum(pad, 0)
    pstart = pend - pad
    Mp = (M_ + 63 * G + 63) // 64 * 64             # static upper bound of the padded length (no host sync)
    pos = torch.arange(Mp, device=dev)
    e = torch.searchsorted(pend, pos, right=True).clamp(max=G - 1)
    j = pos - pstart[e]
    ids = torch.where((pos < pend[-1]) & (j < split[e]), cum[e] + j, torch.full_like(pos, -1)).to(torch.int32)
    a_t = transpose_gather(grad_output, ids, Mp)   # [N, Mp]
    b_t = transpose_gather(original_input, ids, Mp)  # [K, Mp]
    seg = torch.cat([pstart, pend[-1:]]).div(64, rounding_mode="floor").to(torch.int32).contiguous()
    cfg = GemmConfig(bn=256 if K % 256 == 0 else 128, cta_group=2 if N % 256 == 0 else 1, group_m=8, use_tma_store=True)
    args = _C.GemmArgs()
    args.mode = 0
    fill_common(args, N, a_t.data_ptr(), Mp, b_t, grad_weight.data_ptr(), N, K, N, K, Mp, cfg, grad_output.dtype == torch.bfloat16)
    args.c_nbuf, args.c_buf_stride_bytes = G, N * K * grad_weight.element_size()
    args.segk_off, args.segk_n = seg.data_ptr(), G
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), _s()), "td_gemm_launch(wgrad, segmented K)")
    return grad_weight


def _use_tma_gather(x: torch.Tensor) -> bool:
    return x.is_cuda and U.get_bool_env("TD_MOE_TMA_GATHER", _TMA_GATHER_DEFAULT)


_TMA_GATHER_DEFAULT = True      # validated on B200: bit-identical to the staged path (tests/test_ops_gpu.py::test_moe_tma_gather)


def moe_grouped_gemm(x_sorted: torch.Tensor, w: torch.Tensor, routing: SortedRouting, out: Optional[torch.Tensor] = None,
                     config: Optional[GemmConfig] = None) -> torch.Tensor:
    """``y_sorted[i] = x_sorted[i] @ w[expert_of_row(i)].T``;  x_sorted: [capacity, K] (expert-sorted, padded),
    w: [E, N, K] (K-major per expert).  Rows of unused padded tiles are left untouched."""
    cap, K = x_sorted.shape
    E, N, Kw = w.shape
    assert Kw == K and cap == routing.capacity
    out = torch.empty((cap, N), dtype=x_sorted.dtype, device=x_sorted.device) if out is None else out
    if not x_sorted.is_cuda:
        te = routing.tile_expert
        bm = routing.block_m
        for t in range(te.numel()):
            e = int(te[t])
            if e >= 0:
                out[t * bm:(t + 1) * bm] = (x_sorted[t * bm:(t + 1) * bm].float() @ w[e].float().t()).to(out.dtype)
        return out
    cfg = config or GemmConfig(bn=256 if N >= 256 else 128, cta_group=1, group_m=1, use_tma_store=True)
    assert routing.block_m == 128 * cfg.cta_group, "routing block_m must equal the GEMM tile height"
    w2 = w.reshape(E * N, K)
    args = _C.GemmArgs()
    args.mode = 0
    fill_common(args, cap, x_sorted.data_ptr(), x_sorted.stride(0), w2, out.data_ptr(), cap, out.stride(0), cap, N, K, cfg,
                x_sorted.dtype == torch.bfloat16)
    args.tile_expert, args.num_experts = routing.tile_expert.data_ptr(), E
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), _s()), "td_gemm_launch(grouped)")
    return out


def moe_forward_local(x: torch.Tensor, w: torch.Tensor, topk_ids: torch.Tensor, num_experts: Optional[int] = None) -> torch.Tensor:
    """``c[t * topk + j] = x[t] @ w[ids[t, j]].T`` for local tokens (no communication)."""
    E = w.shape[0] if num_experts is None else num_experts
    topk = topk_ids.shape[1]
    if x.is_cuda and topk_ids.numel() >= _PAIRS_FOR_CTA_PAIRS and _use_tma_gather(x):
        return moe_grouped_gemm_presorted(x.contiguous(), w, topk_ids, E, topk)
    r = moe_align_sort(topk_ids, E, 128)
    if _use_tma_gather(x):
        return moe_grouped_gemm_fused(x.contiguous(), w, r, topk, topk_ids.numel())
    xs = gather_rows(x, r, div=topk)
    ys = moe_grouped_gemm(xs, w, r)
    return scatter_rows(ys, r, topk_ids.numel())


def calc_sorted_gather_index(topk_ids: torch.Tensor, num_ranks: int, num_experts: int, block_size: int = 128, rank: int = 0):
    """(sorted (token, k) indices padded per expert to ``block_size``, token counts [num_ranks, num_experts]) -- rows of every expert ordered
    by the all-gather arrival stage of their source rank (reference: allgather_group_gemm.py ``calc_sorted_gather_index`` :86-166)."""
    T = topk_ids.shape[0]
    assert T % num_ranks == 0
    r = moe_align_sort(topk_ids, num_experts, block_size, tokens_per_rank=T // num_ranks, rank=rank, world=num_ranks)
    ids = topk_ids.view(num_ranks, -1).long()
    valid = (ids >= 0) & (ids < num_experts)
    cnt = torch.zeros(num_ranks, num_experts + 1, dtype=torch.int64, device=topk_ids.device)
    cnt.scatter_add_(1, torch.where(valid, ids, torch.full_like(ids, num_experts)), torch.ones_like(ids))
    return r.sorted_ids, cnt[:, :num_experts].to(torch.int32), r


def sort_topk_ids_align_block_size(topk_ids: torch.Tensor, num_experts: int, rank: int, num_ranks: int, num_local_ranks: int = 0,
                                   block_size: int = 128):
    """The routing metadata of the fused AllGather + grouped GEMM in the reference's shape (allgather_group_gemm.py:201):
    ``(sorted_gather_index, expert_idx, tiled_m, segment_start, segment_end, ntiles)`` -- per tile, in execution order (tiles whose
    shards arrive first run first): its expert, its row-block index in the padded sorted layout, and the first / last all-gather stage it
    needs.  The tile table is built on the host from the [ranks, experts] histogram (one small device-to-host copy)."""
    from .tile_swizzle import ag_moe_tile_table
    sorted_ids, cnt, r = calc_sorted_gather_index(topk_ids, num_ranks, num_experts, block_size, rank)
    table = ag_moe_tile_table(cnt.cpu().numpy(), rank, block_size)
    dev = topk_ids.device
    t = torch.from_numpy(table).to(dev) if len(table) else torch.zeros((0, 4), dtype=torch.int32, device=dev)
    first_block = (r.expert_offsets[:-1].to(torch.int64) // block_size).to(dev)
    tiled_m = (first_block[t[:, 0].long()] + t[:, 1].long()).to(torch.int32)
    return sorted_ids, t[:, 0].contiguous(), tiled_m, t[:, 2].contiguous(), t[:, 3].contiguous(), torch.tensor([t.shape[0]], dtype=torch.int32, device=dev)


def run_moe_ag_triton_non_overlap(x_shard: torch.Tensor, weights: torch.Tensor, chosen_experts: torch.Tensor, group=None, **_hints):
    """The NON-overlapped baseline of ``ag_group_gemm`` (reference :901): NCCL all-gather of the token shards, then the local grouped GEMM.
    ``weights``: [E, N_local, K] -> [T * topk, N_local]."""
    group = group or U.get_triton_dist_world()
    W = U.world_size()
    x = torch.empty((x_shard.shape[0] * W, x_shard.shape[1]), dtype=x_shard.dtype, device=x_shard.device)
    torch.distributed.all_gather_into_tensor(x, x_shard.contiguous(), group=group)
    if not x.is_cuda:
        ids = chosen_experts.long()
        return torch.einsum("tk,tjnk->tjn", x.float(), weights.float()[ids]).reshape(ids.numel(), -1).to(x.dtype)
    return moe_forward_local(x, weights, chosen_experts)


# ------------------------------------------------------------------------------------------------------------
# AG + grouped GEMM (TP-MoE up projection)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class MoEAllGatherGroupGEMMContext:
    max_ntokens: int
    N_per_rank: int
    K: int
    num_experts: int
    topk: int
    dtype: torch.dtype
    rank: int
    world_size: int
    ag_ctx: comm.FastAllGatherContext = None
    fused_ctx: object = None        # AllGatherGEMMTensorParallelContext of the single-kernel path (workspace + flags)

    def finalize(self):
        if self.ag_ctx is not None:
            self.ag_ctx.finalize()
            self.ag_ctx = None
        if self.fused_ctx is not None:
            self.fused_ctx.finalize()
            self.fused_ctx = None


def create_ag_group_gemm_context(max_ntokens: int, N_per_rank: int, K: int, num_experts: int, topk: int, dtype: torch.dtype,
                                 rank: Optional[int] = None, world_size: Optional[int] = None, local_world_size=None,
                                 **ref_hints) -> MoEAllGatherGroupGEMMContext:
    U.accept_ref_hints("create_ag_group_gemm_context", ref_hints, ('BLOCK_M', 'BLOCK_N', 'BLOCK_K', 'GROUP_SIZE_M', 'stages', 'num_warps', 'ag_stream', 'group_gemm_stream', 'for_correctness'))
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    ctx = MoEAllGatherGroupGEMMContext(max_ntokens, N_per_rank, K, num_experts, topk, dtype, rank, world_size)
    shard_bytes = (max_ntokens // world_size) * K * torch.empty(0, dtype=dtype).element_size()
    ctx.ag_ctx = comm.create_fast_allgather_context(max(shard_bytes, 1024), rank, world_size, grid_max=64)
    if heap.device.type == "cuda" and world_size > 1 and U.get_bool_env("TD_MOE_AG_FUSED", True):
        from .ag_gemm import create_ag_gemm_context
        ctx.fused_ctx = create_ag_gemm_context(max_ntokens, N_per_rank, K, dtype, rank, world_size)
    return ctx


def _ag_group_gemm_fused(a: torch.Tensor, b: torch.Tensor, ctx: MoEAllGatherGroupGEMMContext, r: SortedRouting, T: int,
                         out: Optional[torch.Tensor], n_comm: int = 16) -> torch.Tensor:
    """AllGather(tokens) + grouped GEMM in ONE kernel: comm CTAs push this rank's tokens into every peer's workspace
    (the ag_gemm protocol), the GEMM producer warp waits -- per lane, for the four rows it fetches -- on the arrival flags
    of the slices that carry them and gathers the rows by TMA ``tile::gather4``; the epilogue scatters to (token, k) order.
    Tiles are ordered by expert and, inside an expert, by the arrival stage of the source rank (moe_align_sort)."""
    g = ctx.fused_ctx
    W, me = ctx.world_size, ctx.rank
    Ms, K = a.shape
    E, N, _ = b.shape
    out = torch.empty((T * ctx.topk, N), dtype=a.dtype, device=a.device) if out is None else out
    a = a.contiguous()
    args = _C.GemmArgs()
    args.mode = 1
    cfg = GemmConfig(bn=256 if N >= 256 else 128, cta_group=1, group_m=1, use_tma_store=False, n_comm_ctas=n_comm)
    ws_buf_bytes = g.max_M * K * a.element_size()
    fill_common(args, 2 * g.max_M, g.workspace.data_ptr(), K, b.reshape(E * N, K), out.data_ptr(), T * ctx.topk, out.stride(0),
                r.capacity, N, K, cfg, a.dtype == torch.bfloat16)
    args.a_nbuf, args.a_buf_stride_bytes = 2, ws_buf_bytes
    args.tile_expert, args.num_experts = r.tile_expert.data_ptr(), E
    args.a_gather, args.a_gather_div, args.a_gather_pad = r.sorted_ids.data_ptr(), ctx.topk, r.pad_id
    args.a_src_rows, args.c_scatter = 2 * g.max_M, r.sorted_ids.data_ptr()
    rk, w, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = rk, w, base, stride, mc
    args.phase = g.phase.data_ptr()
    args.ag_rows_per_rank, args.ag_copy_local, args.ag_skip_wait = Ms, 1, 0
    args.ag_a_local, args.ag_ws, args.ag_ws_buf_bytes = a.data_ptr(), g.workspace.data_ptr(), ws_buf_bytes
    args.ag_flags, args.ag_ready = g.flags.data_ptr(), g.ready.data_ptr()
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), _s()), "td_gemm_launch(ag + grouped, gather4)")
    g.host_phase += 1
    return out


def ag_group_gemm(a: torch.Tensor, b: torch.Tensor, ctx: MoEAllGatherGroupGEMMContext, full_topk_ids: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a: ``[T/W, K]`` local tokens, b: ``[E, N/W, K]`` (K-major expert weights; the reference passes ``[E, K, N/W]`` --
    a ``transpose(1, 2)`` view of this layout is accepted), full_topk_ids: ``[T, topk]`` (already all-gathered)
    -> ``c[T * topk, N/W]`` with row ``t * topk + j = a_full[t] @ b[ids[t, j]]``."""
    W = ctx.world_size
    if b.shape[1] == a.shape[1] and b.shape[2] != a.shape[1]:
        b = b.transpose(1, 2)                      # [E, K, N] -> [E, N, K] view
    if b.stride(2) != 1:
        b = b.contiguous()
    T = full_topk_ids.shape[0]
    tpr = a.shape[0]
    assert tpr * W == T
    if W > 1 and ctx.fused_ctx is not None and a.is_cuda and T <= ctx.fused_ctx.max_M:
        r = moe_align_sort(full_topk_ids, ctx.num_experts, 128, tokens_per_rank=tpr, rank=ctx.rank, world=W)
        return _ag_group_gemm_fused(a, b, ctx, r, T, out)
    # 1. all-gather the tokens (push over NVLink; consumers sorted by arrival stage below)
    if W > 1:
        a_full = comm.fast_allgather(a.contiguous(), ctx.ag_ctx, mode="push").view(T, a.shape[1])
    else:
        a_full = a
    # 2. route: expert-major, and within an expert by arrival stage of the source rank
    r = moe_align_sort(full_topk_ids, ctx.num_experts, 128, tokens_per_rank=tpr, rank=ctx.rank, world=W)
    if _use_tma_gather(a_full):
        return moe_grouped_gemm_fused(a_full, b, r, ctx.topk, T * ctx.topk, out=out)
    xs = gather_rows(a_full, r, div=ctx.topk)
    ys = moe_grouped_gemm(xs, b, r)
    return scatter_rows(ys, r, T * ctx.topk, out=out)


# ------------------------------------------------------------------------------------------------------------
# grouped GEMM + top-k reduce + ReduceScatter / AllReduce (TP-MoE down projection)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class MoEReduceRSContext:
    rank: int
    world_size: int
    max_token_num: int       # T * topk
    hidden_dim: int
    num_experts: int
    topk: int
    dtype: torch.dtype
    ar_ctx: comm.AllReduceContext = None
    # single-kernel path (csrc/gemm_sm100.cuh, mode kMoeRS)
    part: torch.Tensor = None        # symmetric [2, T, N]: my top-k-reduced partial (built by L2 reductions in the epilogue),
                                     # pulled by the owners through the NVSwitch
    flags: torch.Tensor = None       # symmetric int32 [2, chunks, W]: 'rank s finished column chunk c' (phase numbers)
    counter: torch.Tensor = None     # local int32 [2, chunks + 1]: finished CTA tiles per chunk (+ CTAs done zeroing)
    phase: torch.Tensor = None       # local int32 [4]
    n_comm: int = 16

    def finalize(self):
        if self.ar_ctx is not None:
            self.ar_ctx.finalize()
            self.ar_ctx = None
        heap = U.get_heap()
        for t in (self.part, self.flags):
            if t is not None:
                heap.free_tensor(t)
        self.part = self.flags = None


_MRS_MAX_NTILES = 64


def create_moe_rs_context(rank: Optional[int], world_size: Optional[int], local_world_size, max_token_num: int,
                          hidden_dim: int, num_experts: int, topk: int, dtype: torch.dtype, n_chunks_max: int = 8,
                          **ref_hints) -> MoEReduceRSContext:
    U.accept_ref_hints("create_moe_rs_context", ref_hints, ('rs_stream', 'reduction_stream'))
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    ctx = MoEReduceRSContext(rank, world_size, max_token_num, hidden_dim, num_experts, topk, dtype)
    T = max_token_num // topk
    ctx.ar_ctx = comm.create_allreduce_ctx(max(T * hidden_dim * torch.empty(0, dtype=dtype).element_size(), 1024), rank,
                                           world_size, world_size)
    if heap.device.type == "cuda" and world_size > 1 and U.get_bool_env("TD_MOE_RS_FUSED", True):
        ctx.part = heap.tensor((2, T, hidden_dim), dtype)
        ctx.flags = heap.tensor((2, _MRS_MAX_NTILES, world_size, 32), torch.int32)
        ctx.counter = torch.zeros((2, _MRS_MAX_NTILES), dtype=torch.int32, device=heap.device)
        ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
        U.barrier_all_host()
    return ctx


create_moe_ar_context = create_moe_rs_context


def _moe_down_partial(x, w, chosen_experts, expert_weight, ctx):
    """Local part: grouped GEMM over this rank's K-shard + weighted top-k reduction -> partial ``[T, N]``."""
    if w.shape[1] == x.shape[1] and w.shape[2] != x.shape[1]:
        w = w.transpose(1, 2)                      # [E, K/W, N] -> [E, N, K/W]
    if w.stride(2) != 1:
        w = w.contiguous()
    r = None
    if x.is_cuda and chosen_experts.numel() >= _PAIRS_FOR_CTA_PAIRS and _use_tma_gather(x):
        y = moe_grouped_gemm_presorted(x.contiguous(), w, chosen_experts, ctx.num_experts, 1)
    elif _use_tma_gather(x):
        r = moe_align_sort(chosen_experts, ctx.num_experts, 128)
        y = moe_grouped_gemm_fused(x.contiguous(), w, r, 1, chosen_experts.numel())
    else:
        r = moe_align_sort(chosen_experts, ctx.num_experts, 128)
        xs = gather_rows(x, r, div=1)                  # rows of x are already (token, k) pairs
        ys = moe_grouped_gemm(xs, w, r)
        y = scatter_rows(ys, r, chosen_experts.numel())
    return reduce_topk(y, expert_weight, ctx.topk)


def _moe_reduce_fused(x, wk, chosen_experts, expert_weight, ctx: MoEReduceRSContext, allreduce: bool,
                      out: Optional[torch.Tensor] = None, n_comm: Optional[int] = None, bn: Optional[int] = None,
                      chunk_n: int = 0) -> torch.Tensor:
    """Grouped GEMM + weighted top-k reduce + ReduceScatter (or AllReduce) in ONE kernel (csrc/gemm_sm100.cuh, kMoeRS):
    grouped tcgen05 tiles run column-chunk major (chunks shrink towards the end); the epilogue multiplies each row by its
    routing weight and ADDS it to its token's row of this rank's symmetric partial with 16-byte L2 reductions
    (``red.add.v4.bf16x2``) -- the top-k reduce needs no pass and no (token, k) buffer; the CTA that finishes a chunk
    release-flags all ranks; comm CTAs of the same grid pull the cross-rank sum of the rows this rank owns through the NVSwitch
    (``multimem.ld_reduce``, fp32 accumulation).  No reduce_topk kernel, no side stream, no final copy.  The partial is 16-bit:
    with top-k > 2 the L2 adds round more than once per element (use the staged path when that matters).
    Reference: moe_reduce_rs.py:168-246 + 549-619 (two kernels and a stream pair), moe_reduce_ar.py:563."""
    W, topk = ctx.world_size, ctx.topk
    T = chosen_experts.shape[0]
    E, N, K = wk.shape
    assert x.shape == (T * topk, K) and T * topk <= ctx.max_token_num and N == ctx.hidden_dim
    bn = bn or (256 if N % 256 == 0 else 128)
    num_n = (N + bn - 1) // bn
    assert num_n <= _MRS_MAX_NTILES, "moe_reduce_rs: too many n tiles for the flag array"
    n_comm = min(32, n_comm or ctx.n_comm)
    big = T * topk >= _PAIRS_FOR_CTA_PAIRS          # compute bound: pre-sorted rows + CTA pairs; else the gather4 producer
    r = moe_align_sort(chosen_experts, ctx.num_experts, 256 if big else 128)
    assert T > 0
    rows_out = T if allreduce else T // W
    out = torch.empty((rows_out, N), dtype=x.dtype, device=x.device) if out is None else out
    scale = expert_weight.reshape(-1).to(torch.float32).contiguous()
    xc = gather_rows(x.contiguous(), r, div=1) if big else x.contiguous()
    args = _C.GemmArgs()
    args.mode = 4
    fill_common(args, xc.shape[0], xc.data_ptr(), xc.stride(0), wk.reshape(E * N, K), ctx.part.data_ptr(), T, N,
                r.capacity, N, K, GemmConfig(bn, 2 if big else 1, 1, False, 0, n_comm), x.dtype == torch.bfloat16)
    args.tile_expert, args.num_experts = r.tile_expert.data_ptr(), E
    if not big:
        args.a_gather, args.a_gather_div = r.sorted_ids.data_ptr(), 1
        args.a_src_rows = xc.shape[0]
    args.a_gather_pad, args.c_scatter = r.pad_id, r.sorted_ids.data_ptr()
    args.mrs_chunk_n = chunk_n
    rk, w_, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = rk, w_, base, stride, mc
    args.phase = ctx.phase.data_ptr()
    args.rs_stage, args.rs_stage_buf_bytes = ctx.part.data_ptr(), ctx.part[0].numel() * ctx.part.element_size()
    args.rs_flags, args.rs_out, args.rs_ldo = ctx.flags.data_ptr(), out.data_ptr(), out.stride(0)
    args.row_scale, args.mrs_counter, args.mrs_total_padded = scale.data_ptr(), ctx.counter.data_ptr(), r.total_padded.data_ptr()
    args.mrs_T, args.mrs_topk, args.mrs_allreduce = T, topk, int(allreduce)
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), _s()), "td_gemm_launch(moe_reduce_rs)")
    return out


def _fused_ok(x, wk, ctx, T, allreduce):
    W = ctx.world_size
    return (W > 1 and x.is_cuda and ctx.part is not None and _use_tma_gather(x) and wk.is_contiguous() and wk.shape[1] % 8 == 0
            and wk.shape[1] == ctx.hidden_dim and ctx.part.shape[1] >= T and ctx.part.shape[2] == wk.shape[1] and (allreduce or T % W == 0)
            and (wk.shape[1] + 127) // 128 <= _MRS_MAX_NTILES and x.dtype in (torch.bfloat16, torch.float16)
            and U.get_bool_env("TD_MOE_RS_FUSED", True))


def run_moe_reduce_rs(x: torch.Tensor, w: torch.Tensor, chosen_experts: torch.Tensor, expert_weight: torch.Tensor,
                      ctx: MoEReduceRSContext, n_chunks: int = 2, **ref_hints) -> torch.Tensor:
    """x: ``[T*topk, K/W]``, w: ``[E, K/W, N]`` (or K-major ``[E, N, K/W]``), chosen_experts/expert_weight: ``[T, topk]``
    -> ``[T/W, N]`` = reduce_scatter_ranks( sum_j weight[t,j] * (x[t*topk+j] @ w[e_tj]) )."""
    U.accept_ref_hints("run_moe_reduce_rs", ref_hints, ('persistent', 'config'))
    W = ctx.world_size
    wk = w.transpose(1, 2) if (w.shape[1] == x.shape[1] and w.shape[2] != x.shape[1]) else w      # -> [E, N, K/W]
    N = wk.shape[1]
    T = chosen_experts.shape[0]
    if _fused_ok(x, wk, ctx, T, False):
        return _moe_reduce_fused(x, wk, chosen_experts, expert_weight, ctx, False)
    if (W > 1 and x.is_cuda and n_chunks > 1 and _use_tma_gather(x) and wk.is_contiguous() and N % (n_chunks * 128) == 0
            and ((T // W) * (N // n_chunks) * x.element_size()) % 16 == 0):
        return _moe_reduce_rs_chunked(x, wk, chosen_experts, expert_weight, ctx, n_chunks)
    part = _moe_down_partial(x, w, chosen_experts, expert_weight, ctx)
    if W == 1:
        return part
    return comm.reduce_scatter(part.contiguous(), ctx.ar_ctx)


def _moe_reduce_rs_chunked(x, wk, chosen_experts, expert_weight, ctx, n_chunks):
    """The reference's overlap (moe_reduce_rs.py:168-246,549-619): N is split into chunks; while the grouped GEMM of chunk c+1
    runs on the compute stream, chunk c goes through top-k reduce -> reduce-scatter (NVLS pull-reduce) on a side stream."""
    W, topk = ctx.world_size, ctx.topk
    T = chosen_experts.shape[0]
    N = wk.shape[1]
    Nc = N // n_chunks
    big = T * topk >= _PAIRS_FOR_CTA_PAIRS
    r = moe_align_sort(chosen_experts, ctx.num_experts, 256 if big else 128)
    xc = gather_rows(x.contiguous(), r, div=1) if big else x.contiguous()
    out = torch.empty((T // W, N), dtype=x.dtype, device=x.device)
    main = torch.cuda.current_stream()
    if getattr(ctx, "_rs_stream", None) is None:
        ctx._rs_stream = torch.cuda.Stream(priority=-1)
    side = ctx._rs_stream
    side.wait_stream(main)
    for c in range(n_chunks):
        y = moe_grouped_gemm_fused(xc, wk, r, 1, T * topk, n_slice=(c * Nc, Nc), gather=not big)
        part = reduce_topk(y, expert_weight, topk)                     # [T, Nc]
        ev = torch.cuda.Event()
        ev.record(main)
        part.record_stream(side)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            o = comm.reduce_scatter(part, ctx.ar_ctx)
            out[:, c * Nc:(c + 1) * Nc].copy_(o)
    main.wait_stream(side)
    return out


def run_moe_reduce_ar(x, w, chosen_experts, expert_weight, ctx: MoEReduceRSContext, **ref_hints) -> torch.Tensor:
    U.accept_ref_hints("run_moe_reduce_ar", ref_hints, ('n_chunks', 'persistent', 'config'))
    wk = w.transpose(1, 2) if (w.shape[1] == x.shape[1] and w.shape[2] != x.shape[1]) else w      # -> [E, N, K/W]
    if _fused_ok(x, wk, ctx, chosen_experts.shape[0], True):
        return _moe_reduce_fused(x, wk, chosen_experts, expert_weight, ctx, True)
    part = _moe_down_partial(x, w, chosen_experts, expert_weight, ctx)
    if ctx.world_size == 1:
        return part
    return comm.all_reduce(part, None if part.is_cuda else comm.AllReduceMethod.OneShot, ctx.ar_ctx)


def moe_reduce_rs_torch(x, w, chosen_experts, expert_weight, group, world_size, rank):
    """Golden (reference test: test_moe_reduce_rs.py:88-107): masked per-expert matmul, weighted sum, reduce_scatter."""
    if w.shape[1] != x.shape[1]:
        w = w.transpose(1, 2)
    T, topk = chosen_experts.shape
    ids = chosen_experts.reshape(-1).long()
    y = torch.zeros((T * topk, w.shape[2]), dtype=torch.float32, device=x.device)
    for e in range(w.shape[0]):
        m = ids == e
        if m.any():
            y[m] = x[m].float() @ w[e].float()
    part = (y.view(T, topk, -1) * expert_weight.float()[..., None]).sum(1)
    if world_size > 1:
        dist.all_reduce(part, group=group)
    return part[rank * (T // world_size):(rank + 1) * (T // world_size)]


def moe_reduce_ar_torch(x, w, chosen_experts, expert_weight, group, world_size):
    """Golden / non-overlapped baseline of ``run_moe_reduce_ar`` (reference: moe_reduce_ar.py ``run_moe_reduce_ar_triton_non_overlap``):
    masked per-expert matmul, weighted top-k sum, one all-reduce -> [T, N] on every rank."""
    if w.shape[1] != x.shape[1]:
        w = w.transpose(1, 2)
    T, topk = chosen_experts.shape
    ids = chosen_experts.reshape(-1).long()
    y = torch.zeros((T * topk, w.shape[2]), dtype=torch.float32, device=x.device)
    for e in range(w.shape[0]):
        m = ids == e
        if m.any():
            y[m] = x[m].float() @ w[e].float()
    part = (y.view(T, topk, -1) * expert_weight.float()[..., None]).sum(1)
    if world_size > 1:
        dist.all_reduce(part, group=group)
    return part

