"""Sequence / context parallel layers.

* ``SpGQAFlashDecodeAttention`` -- KV cache sharded along the sequence over ranks; each rank runs split-KV flash-decode
  on its shard, the (O, LSE) partials are exchanged with the low-latency all-gather and merged with an LSE-weighted
  combine (reference: layers/nvidia/sp_flash_decode_layer.py:79-185, flash_decode.py:482).
* ``fused_sp_ag_attn_intra_node`` -- context-parallel prefill: q stays sharded, K/V shards are all-gathered over NVLink,
  attention of the local q block runs over the full KV with zig-zag causal balancing
  (reference: kernels/nvidia/sp_ag_attention_intra_node.py:60-522).
* ``UlyssesSPAllToAllLayer`` -- head<->sequence all-to-all before/after attention
  (reference: kernels/nvidia/ulysses_sp_dispatch.py:546-707, layers/nvidia/ulysses_sp_a2a_layer.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .. import utils as U
from ..ops import comm
from ..ops.all_to_all import all_to_all_single_2d, create_all_to_all_single_2d_context
from ..ops.flash_decode import combine_partials, gqa_fwd_batch_decode_partial


class SpGQAFlashDecodeAttention:
    def __init__(self, rank: int, world_size: int, num_q_heads: int, num_kv_heads: int, head_dim: int = 128,
                 max_batch: int = 64, soft_cap: float = 0.0):
        self.rank, self.world_size = rank, world_size
        self.Hq, self.Hkv, self.D, self.soft_cap = num_q_heads, num_kv_heads, head_dim, soft_cap
        shard = max_batch * num_q_heads * (head_dim + 1) * 4            # fp32 O (D) + LSE (1) per head
        self.ag_ctx = comm.create_fast_allgather_context(shard, rank, world_size)

    def forward(self, q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, local_kv_lens: torch.Tensor,
                block_table: Optional[torch.Tensor] = None, sm_scale: Optional[float] = None) -> torch.Tensor:
        """q: [B, Hq, D] (replicated), caches hold THIS rank's slice of every sequence, ``local_kv_lens``: [B]."""
        B = q.shape[0]
        o, lse = gqa_fwd_batch_decode_partial(q, k_cache, v_cache, local_kv_lens, block_table, sm_scale, self.soft_cap)
        if self.world_size == 1:
            return o.to(q.dtype)
        packed = torch.cat([o.reshape(B, self.Hq, self.D), lse.reshape(B, self.Hq, 1)], dim=-1).contiguous()   # fp32
        mode = "push_2d_ll" if packed.numel() * 4 <= 64 * 1024 else "push"
        allp = comm.fast_allgather(packed, self.ag_ctx, mode=mode)                  # [W, B, Hq, D + 1]
        o_parts = allp[..., :self.D].permute(1, 2, 0, 3).contiguous()               # [B, Hq, W, D]
        lse_parts = allp[..., self.D].permute(1, 2, 0).contiguous()                 # [B, Hq, W]
        return combine_partials(o_parts, lse_parts, q.dtype)

    __call__ = forward

    def finalize(self):
        self.ag_ctx.finalize()


# ------------------------------------------------------------------------------------------------------------
# context-parallel prefill (AllGather KV + local attention)
# ------------------------------------------------------------------------------------------------------------
class SPAllGatherAttentionContextIntraNode:
    def __init__(self, max_kv_tokens_per_rank: int, num_kv_heads: int, head_dim: int, dtype: torch.dtype, rank: int, world_size: int):
        self.rank, self.world_size = rank, world_size
        self.shard_elems = max_kv_tokens_per_rank * num_kv_heads * head_dim
        self.ag_k = comm.create_fast_allgather_context(self.shard_elems * torch.empty(0, dtype=dtype).element_size(), rank, world_size, grid_max=64)
        self.ag_v = comm.create_fast_allgather_context(self.shard_elems * torch.empty(0, dtype=dtype).element_size(), rank, world_size, grid_max=64)

    def finalize(self):
        self.ag_k.finalize()
        self.ag_v.finalize()


def create_sp_ag_attention_context_intra_node(max_kv_tokens_per_rank, num_kv_heads, head_dim, dtype, rank=None, world_size=None, **_):
    heap = U.get_heap()
    return SPAllGatherAttentionContextIntraNode(max_kv_tokens_per_rank, num_kv_heads, head_dim, dtype,
                                                heap.rank if rank is None else rank, heap.world if world_size is None else world_size)


def zigzag_positions(S_total: int, world: int, rank: int, device) -> torch.Tensor:
    """Global token positions owned by ``rank`` under zig-zag sharding: chunk r and chunk 2W-1-r of 2W equal chunks,
    which balances causal-attention work across ranks (reference :283,:361)."""
    c = S_total // (2 * world)
    a = torch.arange(rank * c, (rank + 1) * c, device=device)
    b = torch.arange((2 * world - 1 - rank) * c, (2 * world - rank) * c, device=device)
    return torch.cat([a, b])


def fused_sp_ag_attn_intra_node(ctx: SPAllGatherAttentionContextIntraNode, q_shard: torch.Tensor, k_shard: torch.Tensor,
                                v_shard: torch.Tensor, is_causal: bool = True, enable_zig_zag: bool = True,
                                sm_scale: Optional[float] = None) -> torch.Tensor:
    """q/k/v_shard: [S/W, H, D] for one sequence (batch handled by the caller / varlen loop).  Returns this rank's
    attention output ``[S/W, Hq, D]`` over the full (gathered) KV."""
    W, r = ctx.world_size, ctx.rank
    S_local, Hq, D = q_shard.shape
    Hkv = k_shard.shape[1]
    sm_scale = sm_scale or 1.0 / math.sqrt(D)
    if W > 1:
        k_all = comm.fast_allgather(k_shard.contiguous(), ctx.ag_k, mode="push").view(W, S_local, Hkv, D)
        v_all = comm.fast_allgather(v_shard.contiguous(), ctx.ag_v, mode="push").view(W, S_local, Hkv, D)
    else:
        k_all, v_all = k_shard[None], v_shard[None]
    return _sp_attend(q_shard, k_all, v_all, W, r, is_causal, enable_zig_zag, sm_scale)


def _sp_attend(q_shard, k_all, v_all, W, r, is_causal, enable_zig_zag, sm_scale):
    """Attention of this rank's queries [S/W, Hq, D] over the gathered KV [W, S/W, Hkv, D] of ONE sequence."""
    S_local, Hq, D = q_shard.shape
    Hkv = k_all.shape[2]
    S = S_local * W
    dev = q_shard.device
    zz = enable_zig_zag and W > 1
    if (q_shard.is_cuda and D == 128 and q_shard.dtype in (torch.bfloat16, torch.float16) and U.get_bool_env("TD_TCGEN05_PREFILL", True)
            and (not zz or (S_local // 2) % 128 == 0)):
        return _sp_attn_tcgen05(q_shard, k_all, v_all, W, r, is_causal, enable_zig_zag and W > 1, sm_scale)
    if enable_zig_zag and W > 1:
        pos = torch.stack([zigzag_positions(S, W, s, dev) for s in range(W)])      # [W, S_local]
    else:
        pos = torch.arange(S, device=dev).view(W, S_local)
    q_pos = pos[r]
    k_pos = pos.reshape(-1)
    kk = k_all.reshape(S, Hkv, D).repeat_interleave(Hq // Hkv, dim=1)
    vv = v_all.reshape(S, Hkv, D).repeat_interleave(Hq // Hkv, dim=1)
    # library attention (SDPA) over the gathered KV with an explicit position mask (zig-zag order is not monotone)
    mask = (k_pos[None, :] <= q_pos[:, None]) if is_causal else None
    o = torch.nn.functional.scaled_dot_product_attention(q_shard.transpose(0, 1)[None], kk.transpose(0, 1)[None], vv.transpose(0, 1)[None],
                                                         attn_mask=mask[None, None] if mask is not None else None, scale=sm_scale)
    return o[0].transpose(0, 1).contiguous()


def fused_sp_ag_attn_varlen(ctx: SPAllGatherAttentionContextIntraNode, q_shard: torch.Tensor, k_shard: torch.Tensor, v_shard: torch.Tensor,
                            cu_seqlens_q: torch.Tensor, is_causal: bool = True, enable_zig_zag: bool = True,
                            sm_scale: Optional[float] = None) -> torch.Tensor:
    """Packed variable-length batch under context parallelism (reference: sp_ag_attention_intra_node.py:106-183 + :279-360 -- q shard
    lengths in ``cu_seqlens_q``, per-batch KV gathers).  Every rank holds ``len_b / W`` tokens of every sequence b, packed back to back:
    q/k/v_shard [sum_b len_b / W, H, D], ``cu_seqlens_q`` int32 [B + 1] = cumulative SHARD lengths (identical on all ranks).  The whole
    packed K and V shards are gathered ONCE (two all-gather kernels for the batch instead of two per sequence); attention then runs per
    sequence over its slice of the gathered KV.  Returns [sum_b len_b / W, Hq, D]."""
    W, r = ctx.world_size, ctx.rank
    T_local, Hq, D = q_shard.shape
    Hkv = k_shard.shape[1]
    sm_scale = sm_scale or 1.0 / math.sqrt(D)
    if W > 1:
        k_all = comm.fast_allgather(k_shard.contiguous(), ctx.ag_k, mode="push").view(W, T_local, Hkv, D)
        v_all = comm.fast_allgather(v_shard.contiguous(), ctx.ag_v, mode="push").view(W, T_local, Hkv, D)
    else:
        k_all, v_all = k_shard[None], v_shard[None]
    cu = cu_seqlens_q.tolist()          # one host read per call (the reference reads every boundary with .item())
    out = torch.empty_like(q_shard)
    for b in range(len(cu) - 1):
        a, e = cu[b], cu[b + 1]
        if e > a:
            out[a:e] = _sp_attend(q_shard[a:e], k_all[:, a:e], v_all[:, a:e], W, r, is_causal, enable_zig_zag, sm_scale)
    return out


def merge_attention_partials(o1: torch.Tensor, lse1: torch.Tensor, o2: torch.Tensor, lse2: torch.Tensor):
    """Combine two attention results over DISJOINT key sets: ``o = (e^{l1} o1 + e^{l2} o2) / (e^{l1} + e^{l2})``.
    o: [S, H, D] (normalised), lse: [H, S] (-inf where a query saw no key of that set)."""
    m = torch.maximum(lse1, lse2)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    w1, w2 = torch.exp(lse1 - m), torch.exp(lse2 - m)
    den = (w1 + w2).clamp_min(1e-30)
    a1, a2 = (w1 / den).transpose(0, 1)[..., None], (w2 / den).transpose(0, 1)[..., None]
    return (o1.float() * a1 + o2.float() * a2).to(o1.dtype), m + torch.log(den)


def fused_sp_ag_attn_overlapped(ctx: SPAllGatherAttentionContextIntraNode, q_shard: torch.Tensor, k_shard: torch.Tensor,
                                v_shard: torch.Tensor, is_causal: bool = True, enable_zig_zag: bool = True,
                                sm_scale: Optional[float] = None) -> torch.Tensor:
    """Context-parallel attention with the KV all-gather OVERLAPPED with compute (the reference gathers per-batch KV on a side
    stream while attention runs, sp_ag_attention_intra_node.py:106-183, and its inter-node variant starts on arrived chunks,
    sp_ag_attention_inter_node.py:116-190).  Here: the K / V shards are pushed to every peer on a side stream while the tcgen05
    flash kernel already attends to the LOCAL keys (1/W of the work, ready immediately); the remote keys follow in two
    non-causal calls -- under zig-zag sharding every remote chunk is either fully visible or fully invisible to a query
    half, so ``sk`` bounds replace masks -- and the partial results are merged through their log-sum-exps."""
    from ..ops.flash_attn import flash_attn_fwd
    W, r = ctx.world_size, ctx.rank
    S_local, Hq, D = q_shard.shape
    Hkv = k_shard.shape[1]
    sm_scale = sm_scale or 1.0 / math.sqrt(D)
    zz = enable_zig_zag and W > 1 and is_causal
    ok = (q_shard.is_cuda and W > 1 and is_causal and D == 128 and q_shard.dtype in (torch.bfloat16, torch.float16)
          and ((S_local // 2) % 128 == 0 if zz else S_local % 128 == 0))
    if not ok:
        return fused_sp_ag_attn_intra_node(ctx, q_shard, k_shard, v_shard, is_causal, enable_zig_zag, sm_scale)
    main = torch.cuda.current_stream()
    if getattr(ctx, "_ag_stream", None) is None:
        ctx._ag_stream = torch.cuda.Stream(priority=-1)
    side = ctx._ag_stream
    side.wait_stream(main)
    kc, vc = k_shard.contiguous(), v_shard.contiguous()
    with torch.cuda.stream(side):
        k_all = comm.fast_allgather(kc, ctx.ag_k, mode="push").view(W, S_local, Hkv, D)
        v_all = comm.fast_allgather(vc, ctx.ag_v, mode="push").view(W, S_local, Hkv, D)
        ev = torch.cuda.Event()
        ev.record(side)
    k_all.record_stream(main); v_all.record_stream(main)
    dev = q_shard.device
    n_tiles = S_local // 128
    # ---- local keys (no waiting): zig-zag halves A = chunk r, B = chunk 2W-1-r laid out back to back as positions [0, 2c) ----
    tile_pos = (torch.arange(n_tiles, device=dev, dtype=torch.int32) * 128).view(1, -1).contiguous()
    o_loc, lse_loc = flash_attn_fwd(q_shard[None], kc[None], vc[None], causal=True, sm_scale=sm_scale, q_tile_pos=tile_pos, return_lse=True)
    o, lse = o_loc[0], lse_loc[0]
    main.wait_event(ev)
    if zz:
        c = S_local // 2
        # natural order of the remote chunks: chunk i of 2W lives in rank (i < W ? i : 2W-1-i), first / second half
        kc2, vc2 = k_all.reshape(W, 2, c, Hkv, D), v_all.reshape(W, 2, c, Hkv, D)
        nat = [(i, 0) if i < W else (2 * W - 1 - i, 1) for i in range(2 * W)]
        rem = [i for i in range(2 * W) if i not in (r, 2 * W - 1 - r)]
        idx_r = torch.tensor([nat[i][0] for i in rem], device=dev)
        idx_h = torch.tensor([nat[i][1] for i in rem], device=dev)
        k_rem = kc2[idx_r, idx_h].reshape(1, len(rem) * c, Hkv, D)
        v_rem = vc2[idx_r, idx_h].reshape(1, len(rem) * c, Hkv, D)
        vis_a, vis_b = r, 2 * W - 2 - r               # remote chunks below my half A / below my half B (chunk r itself is local)
        parts = ((slice(0, c), vis_a), (slice(c, 2 * c), vis_b))
    else:
        k_rem, v_rem = k_all[:r].reshape(1, r * S_local, Hkv, D), v_all[:r].reshape(1, r * S_local, Hkv, D)
        parts = ((slice(0, S_local), r * S_local // max(S_local, 1)),)
        c = S_local
    out = torch.empty_like(q_shard)
    for sl, n_vis in parts:
        if n_vis <= 0:
            out[sl] = o[sl]
            continue
        o_r, lse_r = flash_attn_fwd(q_shard[None, sl], k_rem, v_rem, causal=False, sm_scale=sm_scale, sk=n_vis * c, return_lse=True)
        out[sl], _ = merge_attention_partials(o[sl], lse[:, sl], o_r[0], lse_r[0])
    return out


def _sp_attn_tcgen05(q_shard, k_all, v_all, W, r, is_causal, zigzag, sm_scale):
    """Attention of the local q block over the gathered KV with the tcgen05 flash kernel.  With zig-zag sharding rank s
    holds chunks s and 2W-1-s of 2W; the gathered KV is put back into natural order (a view permutation + one copy) so
    the causal mask is ``key <= q_pos`` with consecutive positions inside every 128-query tile."""
    from ..ops.flash_attn import flash_attn_fwd
    _, S_local, Hkv, D = k_all.shape
    S = S_local * W
    if zigzag:
        c = S_local // 2
        assert c % 128 == 0, "zig-zag chunks must be multiples of the 128-query tile"
        order = torch.empty(2 * W, dtype=torch.long)
        for s in range(W):
            order[s], order[2 * W - 1 - s] = 2 * s, 2 * s + 1          # natural chunk -> index in the gathered layout
        k_nat = k_all.reshape(2 * W, c, Hkv, D)[order.to(k_all.device)].reshape(1, S, Hkv, D)
        v_nat = v_all.reshape(2 * W, c, Hkv, D)[order.to(v_all.device)].reshape(1, S, Hkv, D)
        starts = torch.cat([torch.arange(r * c, (r + 1) * c, 128), torch.arange((2 * W - 1 - r) * c, (2 * W - r) * c, 128)])
    else:
        k_nat, v_nat = k_all.reshape(1, S, Hkv, D), v_all.reshape(1, S, Hkv, D)
        starts = torch.arange(r * S_local, (r + 1) * S_local, 128)
    tile_pos = starts.to(torch.int32).to(q_shard.device).view(1, -1).contiguous()
    return flash_attn_fwd(q_shard[None], k_nat, v_nat, causal=is_causal, sm_scale=sm_scale, q_tile_pos=tile_pos)[0]


# ------------------------------------------------------------------------------------------------------------
# Ulysses
# ------------------------------------------------------------------------------------------------------------
class UlyssesSPAllToAllLayer:
    """seq-sharded [S/W, H, D]  <->  head-sharded [S, H/W, D] via one all-to-all each way."""

    def __init__(self, max_local_seq: int, num_heads: int, head_dim: int, dtype: torch.dtype, rank: int, world_size: int):
        self.rank, self.world_size = rank, world_size
        self.ctx = create_all_to_all_single_2d_context(max_local_seq, (num_heads // world_size) * head_dim, dtype)

    def pre_attn_a2a(self, x: torch.Tensor) -> torch.Tensor:
        """[S/W, H, D] -> [S, H/W, D]"""
        W = self.world_size
        S_l, H, D = x.shape
        send = x.view(S_l, W, H // W, D).permute(1, 0, 2, 3).reshape(W * S_l, (H // W) * D).contiguous()
        recv = all_to_all_single_2d(self.ctx, send)
        return recv.view(W * S_l, H // W, D)

    def post_attn_a2a(self, x: torch.Tensor) -> torch.Tensor:
        """[S, H/W, D] -> [S/W, H, D]"""
        W = self.world_size
        S, Hl, D = x.shape
        S_l = S // W
        recv = all_to_all_single_2d(self.ctx, x.reshape(S, Hl * D).contiguous())     # block s = rank s's heads for my tokens
        return recv.view(W, S_l, Hl, D).permute(1, 0, 2, 3).reshape(S_l, W * Hl, D).contiguous()

    def pre_attn_qkv_pack_a2a(self, q, k, v):
        return self.pre_attn_a2a(q), self.pre_attn_a2a(k), self.pre_attn_a2a(v)

    def finalize(self):
        self.ctx.finalize()


class UlyssesQKVPackAllToAll:
    """q, k and v in ONE all-to-all (reference ulysses_sp_dispatch.py:39-277 packs the three tensors per destination inside its
    kernel): the heads every destination owns -- Hq/W query heads and Hkv/W key and value heads -- are packed into one
    ``[(Hq + 2 Hkv) / W * D]``-wide row per local token (one copy pass), a single launch of the all-to-all kernel moves them
    (one fence + one flag round instead of three), and the results are views of the receive buffer."""

    def __init__(self, max_local_seq: int, num_q_heads: int, num_kv_heads: int, head_dim: int, dtype: torch.dtype, rank: int, world_size: int):
        assert num_q_heads % world_size == 0 and num_kv_heads % world_size == 0
        self.rank, self.world_size = rank, world_size
        self.Hq, self.Hkv, self.D = num_q_heads, num_kv_heads, head_dim
        self.cols = (num_q_heads + 2 * num_kv_heads) // world_size * head_dim
        self.ctx = create_all_to_all_single_2d_context(max_local_seq, self.cols, dtype)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
        """q ``[S/W, Hq, D]``, k / v ``[S/W, Hkv, D]`` -> q ``[S, Hq/W, D]``, k / v ``[S, Hkv/W, D]``."""
        W, D = self.world_size, self.D
        S_l = q.shape[0]
        hq, hkv = self.Hq // W, self.Hkv // W
        send = torch.empty((W, S_l, self.cols), dtype=q.dtype, device=q.device)
        send[:, :, :hq * D] = q.reshape(S_l, W, hq * D).transpose(0, 1)
        send[:, :, hq * D:(hq + hkv) * D] = k.reshape(S_l, W, hkv * D).transpose(0, 1)
        send[:, :, (hq + hkv) * D:] = v.reshape(S_l, W, hkv * D).transpose(0, 1)
        recv = all_to_all_single_2d(self.ctx, send.view(W * S_l, self.cols))
        r = recv.view(W * S_l, self.cols)
        return (r[:, :hq * D].reshape(W * S_l, hq, D), r[:, hq * D:(hq + hkv) * D].reshape(W * S_l, hkv, D),
                r[:, (hq + hkv) * D:].reshape(W * S_l, hkv, D))

    __call__ = forward

    def finalize(self):
        self.ctx.finalize()
