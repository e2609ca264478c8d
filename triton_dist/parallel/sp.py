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
