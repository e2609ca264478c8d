"""Routing metadata of the throughput-mode EP all-to-all, as stand-alone device-side ops (no host reads on the way).

Reference: kernels/nvidia/ep_a2a.py ``get_dispatch_send_reqs`` (:725), ``get_ag_splits_and_recv_offset_for_dispatch`` (:765) and
ep_a2a_intra_node.py ``get_ag_splits_and_recv_offset_for_dispatch_intra_node`` (:423).  In the reference these are Triton kernels that feed
its dispatch kernel; the dispatch kernels of this framework (csrc/ep_normal_kernels.cu) build their own slot indices with warp
atomics, so these functions exist for callers that want the reference's metadata itself -- the per-(rank, expert, source) receive
offsets that make every expert's tokens contiguous in the receiver's buffer, the per-rank token counts, and the per-node list of tokens
that need to cross the network once.  They are a handful of sort / cumsum / scatter calls on the device plus ONE small all-gather of the
per-expert histogram over the symmetric heap.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import utils as U


def expert_histogram(topk_indices: torch.Tensor, num_experts: int) -> torch.Tensor:
    """int32 [num_experts + 1]: tokens per expert, last entry = dropped (index < 0 or >= num_experts)."""
    idx = topk_indices.reshape(-1).to(torch.int64)
    bad = (idx < 0) | (idx >= num_experts)
    return torch.bincount(torch.where(bad, torch.full_like(idx, num_experts), idx), minlength=num_experts + 1).to(torch.int32)


def get_dispatch_send_reqs(exp_indices: torch.Tensor, experts_per_rank: int, local_world_size: int, nnodes: Optional[int] = None,
                           max_tokens: Optional[int] = None):
    """Which tokens have to be sent to which NODE (once per node, however many of its experts a token picked).

    exp_indices: int [T, topk].  Returns ``(send_reqs [nnodes, max_tokens] int32, counts [nnodes] int32)``: row n lists, in ascending
    order, the tokens with at least one expert on node n, padded with -1."""
    T, _ = exp_indices.shape
    per_node = experts_per_rank * local_world_size
    if nnodes is None:
        nnodes = max(1, U.world_size() // local_world_size)
    max_tokens = T if max_tokens is None else max_tokens
    idx = exp_indices.to(torch.int64)
    node = torch.where(idx >= 0, idx // per_node, torch.full_like(idx, -1))                   # [T, topk]
    hit = (node[None] == torch.arange(nnodes, device=idx.device)[:, None, None]).any(-1)      # [nnodes, T]
    counts = hit.sum(-1).to(torch.int32)
    order = torch.argsort((~hit).to(torch.int8), dim=1, stable=True)                          # hits first, original order kept
    reqs = torch.where(torch.arange(T, device=idx.device)[None] < counts[:, None], order, torch.full_like(order, -1))
    out = torch.full((nnodes, max_tokens), -1, dtype=torch.int32, device=idx.device)
    out[:, :min(T, max_tokens)] = reqs[:, :max_tokens].to(torch.int32)
    return out, counts


def recv_offsets_from_splits(full_splits: torch.Tensor, experts_per_rank: int):
    """full_splits: int [W, E(+1)] (row s = tokens rank s routes to each expert).  Returns

    * ``recv_buf_offset_per_expert`` int32 [W, epr, W]: entry [r, e, s] = first row, in rank r's receive buffer, of the tokens that
      source s sends to r's local expert e -- expert-major, source-minor, so an expert's rows are contiguous (grouped-GEMM layout);
    * ``num_recv_tokens_per_rank`` int32 [W]: rows every rank receives;
    * ``num_input_tokens_per_rank`` int32 [W]: (token, expert) pairs every rank sends (drops excluded)."""
    W = full_splits.shape[0]
    E = experts_per_rank * W
    sp = full_splits[:, :E].to(torch.int64).view(W, W, experts_per_rank)          # [src, dst rank, e]
    by_dst = sp.permute(1, 2, 0).contiguous()                                     # [dst rank, e, src]
    flat = by_dst.view(W, -1)
    offs = (torch.cumsum(flat, 1) - flat).view(W, experts_per_rank, W)
    return offs.to(torch.int32), flat.sum(1).to(torch.int32), sp.sum((1, 2)).to(torch.int32)


def get_ag_splits_and_recv_offset_for_dispatch_intra_node(topk_indices: torch.Tensor, num_experts: int, ag_ctx=None,
                                                          full_scatter_indices: Optional[torch.Tensor] = None):
    """All-gather the per-expert histogram of every rank and derive the receive layout (see ``recv_offsets_from_splits``).

    Returns ``(recv_buf_offset_per_expert, num_recv_tokens_per_rank, num_input_tokens_per_rank, full_splits)``; with
    ``full_scatter_indices`` (int [W * T, topk]: global output row of every (token, expert) pair, all ranks' outputs viewed as one flat
    buffer) a fifth value: the rows of THIS rank's tokens relative to the start of the destination rank's buffer."""
    from .comm import create_fast_allgather_context, fast_allgather
    W, me = U.world_size(), U.rank()
    assert num_experts % W == 0
    epr = num_experts // W
    local = expert_histogram(topk_indices, num_experts)
    own_ctx = ag_ctx is None
    if own_ctx:
        ag_ctx = create_fast_allgather_context(local.numel() * 4)
    full = fast_allgather(local, ag_ctx, mode="push_2d_ll")                       # [W, E + 1]
    if own_ctx:
        U.barrier_all_host()
        ag_ctx.finalize()
    offs, n_recv, n_in = recv_offsets_from_splits(full, epr)
    if full_scatter_indices is None:
        return offs, n_recv, n_in, full
    T, topk = topk_indices.shape
    mine = full_scatter_indices.view(W, T, topk)[me].to(torch.int64)
    rank_start = torch.cumsum(n_recv.to(torch.int64), 0) - n_recv.to(torch.int64)
    dst = (topk_indices.to(torch.int64).clamp(0, num_experts - 1) // epr)
    rel = torch.where((topk_indices >= 0) & (topk_indices < num_experts), mine - rank_start[dst], torch.full_like(mine, -1))
    return offs, n_recv, n_in, full, rel.to(full_scatter_indices.dtype)
