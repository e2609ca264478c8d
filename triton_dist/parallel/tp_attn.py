"""Tensor-parallel GQA attention block (fused wqkv column shard, wo row shard, optional per-head q/k RMSNorm).

API mirrors /root/reference/python/triton_dist/layers/nvidia/tp_attn.py:70-321.  Projections run on the fused
ops (ag_gemm / gemm_rs / gemm_ar); q/k-norm + RoPE + KV append is one CUDA kernel (csrc/elementwise.cu); decode
attention is our split-KV flash-decode (csrc/attention.cu); prefill attention is our tcgen05 flash-attention kernel
(csrc/flash_attn_sm100.cu; the reference calls flash_attn_with_kvcache :242, kept as the fallback for head_dim != 128).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

from .. import utils as U
from ..ops import comm
from ..ops.ag_gemm import ag_gemm, create_ag_gemm_context
from ..ops.elementwise import qk_norm_rope_kv
from ..ops.flash_decode import gqa_fwd_batch_decode
from ..ops.gemm_ar import create_gemm_ar_context_auto, low_latency_gemm_allreduce_op
from ..ops.gemm_rs import create_gemm_rs_context, gemm_rs
from .tp_mlp import _linear, shard_local

_TCGEN05_PREFILL_DEFAULT = True      # validated on B200 (tests/test_flash_attn_gpu.py: 13/13)

try:  # library attention for prefill
    from flash_attn import flash_attn_with_kvcache as _fa_kvcache
except Exception:  # pragma: no cover
    _fa_kvcache = None


def prefill_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_lens: torch.Tensor, q_len: int,
                      sm_scale: float) -> torch.Tensor:
    """q: [B, S, Hq, D]; caches [B, max_len, Hkv, D] already contain the new tokens; causal."""
    B, S, Hq, D = q.shape
    if q.is_cuda and D == 128 and U.get_bool_env("TD_TCGEN05_PREFILL", _TCGEN05_PREFILL_DEFAULT):
        # our tcgen05 flash-attention kernel (csrc/flash_attn_sm100.cu); one launch when all sequences share a length
        from ..ops.flash_attn import flash_attn_fwd
        if U.get_bool_env("TD_FLASH_VARLEN_KERNEL", False) and q.is_contiguous() and k_cache.is_contiguous() and v_cache.is_contiguous():
            # ONE launch for unequal lengths and no host read of kv_lens: the batch as a packed tensor (every query sequence has S rows,
            # every KV slot max_len rows of which kv_lens[b] are used); opt-in until the varlen instantiation has run on hardware
            from ..ops.flash_attn import flash_attn_varlen
            max_len = k_cache.shape[1]
            ar = torch.arange(B + 1, device=q.device, dtype=torch.int32)
            o = flash_attn_varlen(q.view(B * S, Hq, D), k_cache.view(B * max_len, -1, D), v_cache.view(B * max_len, -1, D), ar * S, ar * max_len,
                                  causal=True, sm_scale=sm_scale, max_seqlen_q=S, one_launch=True, seqused_k=kv_lens.to(torch.int32))
            return o.view(B, S, Hq, D)
        lens = kv_lens.tolist()
        if all(x == lens[0] for x in lens):
            return flash_attn_fwd(q, k_cache, v_cache, causal=True, sm_scale=sm_scale, sk=lens[0])
        out = torch.empty_like(q)
        for b, L in enumerate(lens):
            flash_attn_fwd(q[b:b + 1], k_cache[b:b + 1], v_cache[b:b + 1], causal=True, sm_scale=sm_scale, sk=L, out=out[b:b + 1])
        return out
    if q.is_cuda and _fa_kvcache is not None:
        return _fa_kvcache(q, k_cache, v_cache, cache_seqlens=kv_lens.to(torch.int32), softmax_scale=sm_scale, causal=True)
    outs = []
    for b in range(B):
        L = int(kv_lens[b])
        k, v = k_cache[b, :L], v_cache[b, :L]
        G = Hq // k.shape[1]
        kk, vv = k.repeat_interleave(G, dim=1), v.repeat_interleave(G, dim=1)
        s = torch.einsum("shd,lhd->hsl", q[b].float(), kk.float()) * sm_scale
        qpos = torch.arange(L - S, L, device=q.device)[:, None]
        mask = torch.arange(L, device=q.device)[None, :] <= qpos
        s = s.masked_fill(~mask[None], float("-inf"))
        outs.append(torch.einsum("hsl,lhd->shd", torch.softmax(s, -1), vv.float()).to(q.dtype))
    return torch.stack(outs)


class TP_Attn:
    def __init__(self, rank: int = 0, world_size: int = 8, group=None):
        self.rank, self.world_size, self.group = rank, world_size, group
        self.wqkv = self.wo = self.q_norm_w = self.k_norm_w = None
        self.ag_ctx = self.rs_ctx = self.ar_ctx = self.gemm_ar_ctx = None
        self.ar_method = comm.AllReduceMethod.Unknown

    # ---- parameters -------------------------------------------------------------------------------------
    def _init_parameters(self, attn, verbose: bool = False):
        """``attn``: HF-style module with q_proj/k_proj/v_proj/o_proj (+ optional q_norm/k_norm) and a config."""
        dev = U.current_device()
        W, r = self.world_size, self.rank
        q = shard_local(attn.q_proj.weight.detach(), W, 0, r)
        k = shard_local(attn.k_proj.weight.detach(), W, 0, r)
        v = shard_local(attn.v_proj.weight.detach(), W, 0, r)
        wqkv = torch.cat((q, k, v), dim=0).to(dev)
        wo = shard_local(attn.o_proj.weight.detach(), W, 1, r).to(dev)
        qn = getattr(attn, "q_norm", None)
        kn = getattr(attn, "k_norm", None)
        cfg = attn.config
        self._init_parameters_from_shards(wqkv, wo, qn.weight.detach().to(dev) if qn is not None else None,
                                          kn.weight.detach().to(dev) if kn is not None else None,
                                          cfg.num_attention_heads, cfg.num_key_value_heads,
                                          getattr(cfg, "head_dim", cfg.hidden_size // cfg.num_attention_heads),
                                          getattr(cfg, "rms_norm_eps", 1e-6), getattr(cfg, "rope_theta", 1e6))

    def _init_parameters_from_shards(self, wqkv, wo, q_norm_w, k_norm_w, num_heads, num_kv_heads, head_dim, eps, rope_theta):
        self.wqkv, self.wo, self.q_norm_w, self.k_norm_w = wqkv, wo, q_norm_w, k_norm_w
        self.Hq, self.Hkv, self.D = num_heads // self.world_size, max(1, num_kv_heads // self.world_size), head_dim
        self.eps, self.rope_theta = eps, rope_theta
        self.sm_scale = 1.0 / math.sqrt(head_dim)
        self.dtype = wqkv.dtype
        self.hidden = wqkv.shape[1]

    # ---- contexts -----------------------------------------------------------------------------------------
    def _init_ctx(self, max_M: int, ag_intranode_stream=None, ag_internode_stream=None):
        self.ag_ctx = create_ag_gemm_context(max_M, self.wqkv.shape[0], self.hidden, self.dtype, self.rank, self.world_size)
        self.rs_ctx = create_gemm_rs_context(max_M, self.hidden, self.rank, self.world_size, self.world_size, self.dtype)
        U.barrier_all_host()

    def _init_AR_ctx(self, max_M: int, method=comm.AllReduceMethod.Unknown, dtype=torch.bfloat16):
        self.ar_method = method
        self.ar_ctx = comm.create_allreduce_ctx(max_M * self.hidden * torch.empty(0, dtype=dtype).element_size(), self.rank,
                                                self.world_size, self.world_size)

    def _init_gemm_ar_ctx(self, max_M: int, dtype=torch.bfloat16):
        self.gemm_ar_ctx = create_gemm_ar_context_auto(self.rank, self.world_size, max_M, self.hidden, dtype)

    def finalize(self):
        for c in (self.ag_ctx, self.rs_ctx, self.ar_ctx, self.gemm_ar_ctx):
            if c is not None:
                c.finalize()
        self.ag_ctx = self.rs_ctx = self.ar_ctx = self.gemm_ar_ctx = None

    # ---- shared middle: norm + rope + cache + attention -------------------------------------------------
    def _attn_core(self, qkv: torch.Tensor, position_ids: torch.Tensor, kv_cache, layer_idx: int, bsz: int, q_len: int):
        """qkv: [bsz*q_len, (Hq+2Hkv)*D] -> [bsz*q_len, Hq*D]."""
        k_cache, v_cache = kv_cache.layer(layer_idx)
        T = bsz * q_len
        pos = position_ids.reshape(-1).to(torch.int32)
        bidx = kv_cache.batch_index(bsz, q_len)
        q = qk_norm_rope_kv(qkv, k_cache, v_cache, pos, bidx, self.Hq, self.Hkv, self.q_norm_w, self.k_norm_w, self.eps,
                            self.rope_theta)
        kv_lens = kv_cache.kv_lens_after(q_len)
        if q_len == 1:
            o = gqa_fwd_batch_decode(q.view(bsz, self.Hq, self.D), k_cache, v_cache, kv_lens, sm_scale=self.sm_scale)
        else:
            o = prefill_attention(q.view(bsz, q_len, self.Hq, self.D), k_cache, v_cache, kv_lens, q_len, self.sm_scale)
        return o.reshape(T, self.Hq * self.D)

    # ---- forwards -----------------------------------------------------------------------------------------
    @torch.inference_mode()
    def torch_fwd(self, x, position_ids, kv_cache, layer_idx: int):
        bsz, q_len, H = x.shape
        qkv = torch.nn.functional.linear(x.view(-1, H), self.wqkv)
        o = self._attn_core(qkv, position_ids, kv_cache, layer_idx, bsz, q_len)
        out = torch.nn.functional.linear(o, self.wo)
        if self.world_size > 1:
            dist.all_reduce(out, group=self.group)
        return out.view(bsz, q_len, H)

    @torch.inference_mode()
    def dist_triton_fwd(self, x, position_ids, kv_cache, layer_idx: int):
        """``x``: batch-sharded ``[bsz/W, q_len, H]`` -> same shape.  ag_gemm -> attention -> gemm_rs."""
        b_local, q_len, H = x.shape
        bsz = b_local * self.world_size
        qkv = ag_gemm(x.reshape(-1, H), self.wqkv.t(), self.ag_ctx)
        o = self._attn_core(qkv, position_ids, kv_cache, layer_idx, bsz, q_len)
        out = gemm_rs(o, self.wo.t(), self.rs_ctx)
        return out.view(b_local, q_len, H)

    @torch.inference_mode()
    def dist_triton_AR_fwd(self, x, position_ids, kv_cache, layer_idx: int):
        bsz, q_len, H = x.shape
        qkv = _linear(x.reshape(-1, H), self.wqkv)
        o = self._attn_core(qkv, position_ids, kv_cache, layer_idx, bsz, q_len)
        out = _linear(o, self.wo)
        if self.world_size > 1:
            out = comm.all_reduce(out.contiguous(), self.ar_method, self.ar_ctx)
        return out.view(bsz, q_len, H)

    @torch.inference_mode()
    def dist_triton_gemm_ar_fwd(self, x, position_ids, kv_cache, layer_idx: int):
        bsz, q_len, H = x.shape
        qkv = _linear(x.reshape(-1, H), self.wqkv)
        o = self._attn_core(qkv, position_ids, kv_cache, layer_idx, bsz, q_len)
        out = low_latency_gemm_allreduce_op(self.gemm_ar_ctx, o, self.wo)
        return out.view(bsz, q_len, H)

    def fwd(self, *a, **k):
        raise NotImplementedError("use torch_fwd / dist_triton_fwd / dist_triton_AR_fwd / dist_triton_gemm_ar_fwd")


def layer_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """RMSNorm over the last dimension with weight ``w`` (the helper the reference's tp_attn.py exposes under this name for its
    q / k norms): the CUDA rmsnorm kernel on a GPU, fp32 math otherwise."""
    from ..ops.elementwise import rmsnorm
    shp = x.shape
    return rmsnorm(x.reshape(-1, shp[-1]), w, eps).reshape(shp)
