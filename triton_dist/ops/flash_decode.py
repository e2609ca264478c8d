"""GQA flash-decode: split-KV partials + LSE combine; single GPU or KV sharded across ranks.

Reference: ``gqa_fwd_batch_decode*`` (/root/reference/python/triton_dist/kernels/nvidia/flash_decode.py:763-1132)
and ``SpGQAFlashDecodeAttention`` (layers/nvidia/sp_flash_decode_layer.py:79-185).  Kernels: csrc/attention.cu."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from .. import _C

c_void_p, c_ll, c_int, c_double = C.c_void_p, C.c_longlong, C.c_int, C.c_double


class _DecodeArgs(C.Structure):
    _fields_ = [("q", c_void_p), ("k_cache", c_void_p), ("v_cache", c_void_p), ("kv_lens", c_void_p),
                ("block_table", c_void_p), ("o_part", c_void_p), ("lse_part", c_void_p),
                ("B", c_ll), ("Hq", c_ll), ("Hkv", c_ll), ("S", c_ll), ("max_len", c_ll), ("page_size", c_ll),
                ("max_pages", c_ll), ("is_bf16", c_ll), ("sm_scale", c_double), ("soft_cap", c_double)]


_C.register("td_flash_decode_split", c_int, [C.POINTER(_DecodeArgs), c_void_p])
_C.register("td_flash_decode_combine", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p])


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def default_num_splits(B: int, Hkv: int, max_len: int, sms: int = 148) -> int:
    """Enough CTAs to cover the SMs, but at least ~256 positions per split."""
    want = max(1, (2 * sms) // max(1, B * Hkv))
    return int(max(1, min(want, (max_len + 255) // 256, 64)))


def _decode_reference(q, k_cache, v_cache, kv_lens, sm_scale, block_table=None, page_size=0, soft_cap: float = 0.0):
    B, Hq, D = q.shape
    outs, lses = [], []
    for b in range(B):
        L = int(kv_lens[b])
        if block_table is not None:
            pages = block_table[b, :(L + page_size - 1) // page_size].long()
            k = k_cache[pages].reshape(-1, k_cache.shape[-2], D)[:L]
            v = v_cache[pages].reshape(-1, v_cache.shape[-2], D)[:L]
        else:
            k, v = k_cache[b, :L], v_cache[b, :L]
        Hkv = k.shape[1]
        G = Hq // Hkv
        kk = k.float().repeat_interleave(G, dim=1)          # [L, Hq, D]
        vv = v.float().repeat_interleave(G, dim=1)
        s = torch.einsum("hd,lhd->hl", q[b].float(), kk) * sm_scale
        if soft_cap and soft_cap > 0:
            s = soft_cap * torch.tanh(s / soft_cap)
        lse = torch.logsumexp(s, dim=-1)
        p = torch.softmax(s, dim=-1)
        outs.append(torch.einsum("hl,lhd->hd", p, vv))
        lses.append(lse)
    return torch.stack(outs), torch.stack(lses)


def gqa_fwd_batch_decode_partial(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_lens: torch.Tensor,
                                 block_table: Optional[torch.Tensor] = None, sm_scale: Optional[float] = None,
                                 soft_cap: float = 0.0, num_splits: Optional[int] = None):
    """Local attention over this rank's KV: returns normalised ``O`` (fp32 [B,Hq,D]) and ``LSE`` ([B,Hq])."""
    B, Hq, D = q.shape
    sm_scale = sm_scale if sm_scale is not None else 1.0 / math.sqrt(D)
    if not q.is_cuda or D != 128:
        page = k_cache.shape[1] if block_table is not None else 0
        return _decode_reference(q, k_cache, v_cache, kv_lens, sm_scale, block_table, page, soft_cap)
    Hkv = k_cache.shape[-2]
    if block_table is not None:
        page_size, max_pages, max_len = k_cache.shape[1], block_table.shape[1], 0
        cap = page_size * max_pages
    else:
        page_size, max_pages, max_len = 0, 0, k_cache.shape[1]
        cap = max_len
    S = num_splits or default_num_splits(B, Hkv, cap)
    o_part = torch.empty((B, Hq, S, D), dtype=torch.float32, device=q.device)
    lse_part = torch.empty((B, Hq, S), dtype=torch.float32, device=q.device)
    a = _DecodeArgs()
    a.q, a.k_cache, a.v_cache = q.contiguous().data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
    a.kv_lens = kv_lens.data_ptr()
    a.block_table = block_table.data_ptr() if block_table is not None else None
    a.o_part, a.lse_part = o_part.data_ptr(), lse_part.data_ptr()
    a.B, a.Hq, a.Hkv, a.S, a.max_len, a.page_size, a.max_pages = B, Hq, Hkv, S, max_len, page_size, max_pages
    a.is_bf16, a.sm_scale, a.soft_cap = int(q.dtype == torch.bfloat16), sm_scale, soft_cap
    lib = _C.cuda_lib()
    _C.check(lib.td_flash_decode_split(C.byref(a), _s()), "td_flash_decode_split")
    o = torch.empty((B, Hq, D), dtype=torch.float32, device=q.device)
    lse = torch.empty((B, Hq), dtype=torch.float32, device=q.device)
    _C.check(lib.td_flash_decode_combine(None, o.data_ptr(), lse.data_ptr(), o_part.data_ptr(), lse_part.data_ptr(), B * Hq, S, 0, _s()),
             "td_flash_decode_combine")
    return o, lse


def combine_partials(o_parts: torch.Tensor, lse_parts: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """Merge ``n`` normalised partials: o_parts [B,Hq,n,D] fp32, lse_parts [B,Hq,n] -> [B,Hq,D] (LSE weighting)."""
    B, Hq, n, D = o_parts.shape
    if not o_parts.is_cuda or D != 128 or out_dtype not in (torch.bfloat16, torch.float16):
        w = torch.softmax(lse_parts.float(), dim=-1)
        w = torch.nan_to_num(w, nan=0.0)
        return (o_parts.float() * w[..., None]).sum(2).to(out_dtype)
    out = torch.empty((B, Hq, D), dtype=out_dtype, device=o_parts.device)
    _C.check(_C.cuda_lib().td_flash_decode_combine(out.data_ptr(), None, None, o_parts.contiguous().data_ptr(),
                                                   lse_parts.contiguous().data_ptr(), B * Hq, n, int(out_dtype == torch.bfloat16), _s()),
             "td_flash_decode_combine")
    return out


def gqa_fwd_batch_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_lens: torch.Tensor,
                         block_table: Optional[torch.Tensor] = None, sm_scale: Optional[float] = None,
                         soft_cap: float = 0.0, num_splits: Optional[int] = None) -> torch.Tensor:
    """Single-rank decode attention: q [B,Hq,128] against a (paged) KV cache -> [B,Hq,128]."""
    o, lse = gqa_fwd_batch_decode_partial(q, k_cache, v_cache, kv_lens, block_table, sm_scale, soft_cap, num_splits)
    return o.to(q.dtype)


gqa_fwd_batch_decode_persistent = gqa_fwd_batch_decode
gqa_fwd_batch_decode_intra_rank = gqa_fwd_batch_decode_partial
# the reference's *_aot entry points call kernels compiled ahead of time (flash_decode.py:763-1132); every kernel here is nvcc-built AOT
gqa_fwd_batch_decode_aot = gqa_fwd_batch_decode
gqa_fwd_batch_decode_persistent_aot = gqa_fwd_batch_decode_persistent
gqa_fwd_batch_decode_intra_rank_aot = gqa_fwd_batch_decode_intra_rank

