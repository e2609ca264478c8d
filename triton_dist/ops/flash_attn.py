"""Flash-attention forward (prefill) on tcgen05: ``csrc/flash_attn_sm100.cu``.

Replaces the library call the reference makes for prefill (``flash_attn_with_kvcache`` in layers/nvidia/tp_attn.py:213-247)
and its Triton flash kernel for context-parallel prefill (kernels/nvidia/sp_ag_attention_intra_node.py:257-427):
GQA, causal with an arbitrary per-tile query position (zig-zag sharding), key-length bound, bf16 / fp16, head_dim 128.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from .. import _C

c_void_p, c_ll, c_int, c_double = C.c_void_p, C.c_longlong, C.c_int, C.c_double


class _FlashArgs(C.Structure):
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p), ("lse", c_void_p), ("q_tile_pos", c_void_p),
                ("B", c_ll), ("Sq", c_ll), ("Sk", c_ll), ("Hq", c_ll), ("Hkv", c_ll), ("D", c_ll),
                ("q_stride_b", c_ll), ("q_stride_s", c_ll), ("q_stride_h", c_ll),
                ("k_stride_b", c_ll), ("k_stride_s", c_ll), ("k_stride_h", c_ll),
                ("v_stride_b", c_ll), ("v_stride_s", c_ll), ("v_stride_h", c_ll),
                ("o_stride_b", c_ll), ("o_stride_s", c_ll), ("o_stride_h", c_ll),
                ("sm_scale", c_double), ("causal", c_ll), ("is_bf16", c_ll), ("block_n", c_ll),
                ("cu_q", c_void_p), ("cu_k", c_void_p), ("max_sq", c_ll), ("seqused_k", c_void_p)]


_C.register("td_flash_attn_fwd", c_int, [C.POINTER(_FlashArgs), c_void_p])

Q_TILE = 128


def flash_attn_reference(q, k, v, causal=True, sm_scale=None, q_pos: Optional[torch.Tensor] = None, sk: Optional[int] = None):
    """fp32 golden: q [B, Sq, Hq, D], k/v [B, Sk, Hkv, D]; ``q_pos`` [B, Sq] = position of each query among the keys
    (default: the queries are the last Sq positions).  Returns (out [B, Sq, Hq, D] fp32, lse [B, Hq, Sq])."""
    B, Sq, Hq, D = q.shape
    Sk = k.shape[1] if sk is None else sk
    G = Hq // k.shape[2]
    sm_scale = sm_scale or 1.0 / math.sqrt(D)
    kk = k[:, :Sk].float().repeat_interleave(G, dim=2)
    vv = v[:, :Sk].float().repeat_interleave(G, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), kk) * sm_scale
    if causal:
        if q_pos is None:
            q_pos = torch.arange(Sk - Sq, Sk, device=q.device)[None].expand(B, Sq)
        mask = torch.arange(Sk, device=q.device)[None, None, :] <= q_pos[:, :, None]
        s = s.masked_fill(~mask[:, None], float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    out = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), vv)
    return out, lse


def flash_attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True, sm_scale: Optional[float] = None,
                   q_tile_pos: Optional[torch.Tensor] = None, sk: Optional[int] = None, return_lse: bool = False,
                   out: Optional[torch.Tensor] = None, block_n: int = 64, tmem_operands: bool = False, v2: Optional[bool] = None, v3: bool = False):
    """q: [B, Sq, Hq, 128]; k, v: [B, Sk, Hkv, 128] (any batch/seq/head strides, head dim contiguous).

    ``q_tile_pos``: int32 [B, ceil(Sq / 128)] -- KV position of the first query of every 128-query tile (queries inside
    a tile are consecutive); default ``Sk - Sq + tile * 128`` (the queries are the newest tokens).  ``sk`` bounds the
    keys that exist (a KV cache longer than the sequence).  ``block_n``: keys per pipeline step -- 64 runs two CTAs per
    SM (softmax of one under the MMAs of the other), 128 one CTA per SM with twice the tile.  ``tmem_operands``: 128-key
    tiles with Q and P held in TMEM (both MMAs read only K / V from shared memory).  ``v2``: TMEM operands + two softmax
    warpgroups splitting the columns + 3-stage K/V ring + longest-first CTA order (default).  ``v3`` (opt-in, not yet validated on
    hardware): two query tiles per CTA with one softmax warpgroup each."""
    B, Sq, Hq, D = q.shape
    Sk = k.shape[1] if sk is None else int(sk)
    sm_scale = sm_scale or 1.0 / math.sqrt(D)
    if not q.is_cuda:
        q_pos = None
        if q_tile_pos is not None:
            q_pos = (q_tile_pos.long()[:, :, None] + torch.arange(Q_TILE)[None, None, :]).reshape(B, -1)[:, :Sq]
        o, lse = flash_attn_reference(q, k, v, causal, sm_scale, q_pos, Sk)
        o = o.to(q.dtype)
        if out is not None:
            out.copy_(o)
            o = out
        return (o, lse) if return_lse else o
    assert D == 128 and q.dtype in (torch.bfloat16, torch.float16) and k.dtype == q.dtype == v.dtype
    assert q.stride(3) == 1 and k.stride(3) == 1 and v.stride(3) == 1
    if out is None:
        out = torch.empty((B, Sq, Hq, D), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, Hq, Sq), dtype=torch.float32, device=q.device) if return_lse else None
    a = _FlashArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.lse = lse.data_ptr() if lse is not None else None
    if q_tile_pos is not None:
        assert q_tile_pos.dtype == torch.int32 and q_tile_pos.is_contiguous() and q_tile_pos.numel() == B * ((Sq + Q_TILE - 1) // Q_TILE)
        a.q_tile_pos = q_tile_pos.data_ptr()
    a.B, a.Sq, a.Sk, a.Hq, a.Hkv, a.D = B, Sq, Sk, Hq, k.shape[2], D
    a.q_stride_b, a.q_stride_s, a.q_stride_h = q.stride(0), q.stride(1), q.stride(2)
    a.k_stride_b, a.k_stride_s, a.k_stride_h = k.stride(0), k.stride(1), k.stride(2)
    a.v_stride_b, a.v_stride_s, a.v_stride_h = v.stride(0), v.stride(1), v.stride(2)
    a.o_stride_b, a.o_stride_s, a.o_stride_h = out.stride(0), out.stride(1), out.stride(2)
    a.sm_scale, a.causal, a.is_bf16 = float(sm_scale), int(causal), int(q.dtype == torch.bfloat16)
    if v2 is None:            # default: the v2 kernel unless another variant was asked for explicitly
        v2 = (block_n == 64 and not tmem_operands and not v3)
    a.block_n = 131 if v3 else 130 if v2 else 129 if tmem_operands else (128 if block_n == 128 else 64)
    _C.check(_C.cuda_lib().td_flash_attn_fwd(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_flash_attn_fwd")
    return (out, lse) if return_lse else out


def flash_attn_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens_q: torch.Tensor, cu_seqlens_k: torch.Tensor,
                      causal: bool = True, sm_scale: Optional[float] = None, max_seqlen_q: Optional[int] = None,
                      return_lse: bool = False, one_launch: Optional[bool] = None, seqused_k: Optional[torch.Tensor] = None):
    """Packed variable-length batch (``flash_attn_varlen_func`` semantics): q [Tq, Hq, D], k / v [Tk, Hkv, D], ``cu_seqlens_*`` int32
    [B + 1].  Causal masks are bottom-right aligned per sequence (query i of a sequence sees its keys up to ``Sk - Sq + i``).

    ``one_launch=True`` (or ``TD_FLASH_VARLEN_KERNEL=1``): ONE launch of the varlen instantiation of the v2 kernel -- the cumulative
    lengths stay on the device (no host sync; ``max_seqlen_q`` bounds the grid, default: Tq), CTAs of tiles a sequence does not have
    exit immediately.  Default: one launch per sequence (the host reads the cumulative lengths once); the one-launch kernel compiles
    but has not run on hardware yet.  With ``return_lse`` the LSE comes back as [Hq, Tq].  ``seqused_k`` (int32 [B]): only the first
    ``seqused_k[b]`` keys of slot ``[cu_seqlens_k[b], cu_seqlens_k[b+1])`` exist -- a padded KV cache viewed as a packed tensor."""
    import os
    Tq, Hq, D = q.shape
    if one_launch is None:
        one_launch = os.environ.get("TD_FLASH_VARLEN_KERNEL", "0") == "1"
    B = cu_seqlens_q.numel() - 1
    if q.is_cuda and one_launch:
        assert D == 128 and q.dtype in (torch.bfloat16, torch.float16) and k.dtype == q.dtype == v.dtype
        assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
        assert cu_seqlens_q.dtype == torch.int32 and cu_seqlens_k.dtype == torch.int32 and cu_seqlens_q.is_cuda and cu_seqlens_k.is_cuda
        out = torch.empty((Tq, Hq, D), dtype=q.dtype, device=q.device)
        lse = torch.empty((Hq, Tq), dtype=torch.float32, device=q.device) if return_lse else None
        a = _FlashArgs()
        a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
        a.lse = lse.data_ptr() if lse is not None else None
        a.B, a.Sq, a.Sk, a.Hq, a.Hkv, a.D = B, Tq, k.shape[0], Hq, k.shape[1], D
        a.q_stride_b, a.q_stride_s, a.q_stride_h = Tq * q.stride(0), q.stride(0), q.stride(1)
        a.k_stride_b, a.k_stride_s, a.k_stride_h = k.shape[0] * k.stride(0), k.stride(0), k.stride(1)
        a.v_stride_b, a.v_stride_s, a.v_stride_h = v.shape[0] * v.stride(0), v.stride(0), v.stride(1)
        a.o_stride_b, a.o_stride_s, a.o_stride_h = Tq * out.stride(0), out.stride(0), out.stride(1)
        a.sm_scale, a.causal, a.is_bf16 = float(sm_scale or 1.0 / math.sqrt(D)), int(causal), int(q.dtype == torch.bfloat16)
        a.block_n = 130
        a.cu_q, a.cu_k = cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr()
        a.max_sq = int(max_seqlen_q) if max_seqlen_q else Tq
        if seqused_k is not None:
            assert seqused_k.dtype == torch.int32 and seqused_k.is_cuda and seqused_k.numel() == B
            a.seqused_k = seqused_k.data_ptr()
        _C.check(_C.cuda_lib().td_flash_attn_fwd(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_flash_attn_fwd(varlen)")
        return (out, lse) if return_lse else out
    out = torch.empty_like(q)
    lse = torch.full((Hq, Tq), float("-inf"), dtype=torch.float32, device=q.device) if return_lse else None
    cq, ck = cu_seqlens_q.tolist(), cu_seqlens_k.tolist()
    if seqused_k is not None:
        ck_end = [ck[i] + int(n) for i, n in enumerate(seqused_k.tolist())]
    else:
        ck_end = ck[1:]
    for i in range(len(cq) - 1):
        if cq[i + 1] > cq[i] and ck_end[i] > ck[i]:
            r = flash_attn_fwd(q[None, cq[i]:cq[i + 1]], k[None, ck[i]:ck_end[i]], v[None, ck[i]:ck_end[i]], causal, sm_scale,
                               out=out[None, cq[i]:cq[i + 1]], return_lse=return_lse)
            if return_lse:
                lse[:, cq[i]:cq[i + 1]] = r[1][0]
    return (out, lse) if return_lse else out
