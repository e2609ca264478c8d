"""Qwen3-MoE with tensor-parallel MoE blocks (reference: /root/reference/python/triton_dist/models/qwen_moe.py)."""
from __future__ import annotations

import torch

from ..parallel.tp_moe import TP_MoE
from .config import ArchConfig
from .dense import DenseLLM, DenseLLMLayer, _rand


class _MoEAsMLP:
    """Adapter so DenseLLMLayer can drive a TP_MoE block through the MLP call sites."""

    def __init__(self, moe: TP_MoE):
        self.moe = moe
        self.ag_ctx = self.rs_ctx = self.ar_ctx = self.gemm_ar_ctx = None
        self.ar_method = None

    def torch_fwd(self, x):
        return self.moe.torch_fwd(x)

    def dist_triton_fwd(self, x):
        return self.moe.dist_triton_fwd(x)

    def dist_triton_AR_fwd(self, x):
        return self.moe.dist_triton_AR_fwd(x)

    dist_triton_gemm_ar_fwd = dist_triton_AR_fwd

    def _init_ctx(self, max_M, *a, **k):
        self.moe._init_ctx(max_M)

    def finalize(self):
        self.moe.finalize()


class Qwen3MoELayer(DenseLLMLayer):
    def init_random(self, arch: ArchConfig, dtype, device, seed: int, rank: int, world: int):
        import dataclasses
        # attention + norms exactly like the dense layer (intermediate size irrelevant there)
        super().init_random(dataclasses.replace(arch, intermediate_size=8 * world), dtype, device, seed, rank, world)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + self.layer_idx * 131 + 7 * rank + 3)
        gr = torch.Generator(device=device)
        gr.manual_seed(seed * 104729 + self.layer_idx)
        E, H, I = arch.num_experts, arch.hidden_size, arch.moe_intermediate_size
        moe = TP_MoE(rank, world, self.attn.group)
        moe._init_parameters_from_shards(_rand((E, H), dtype, device, gr, 0.2), _rand((E, 2 * I // world, H), dtype, device, g),
                                         _rand((E, H, I // world), dtype, device, g), arch.num_experts_per_tok, arch.norm_topk_prob)
        self.mlp = _MoEAsMLP(moe)


class Qwen3MoE(DenseLLM):
    def __init__(self, model_config, group=None):
        self._layer_cls = Qwen3MoELayer
        super().__init__(model_config, group)

    def init_triton_dist_ctx(self, max_M: int = 4096):
        l0 = self.layers[0]
        l0.attn._init_ctx(max_M)
        for l in self.layers:
            l.attn.ag_ctx, l.attn.rs_ctx = l0.attn.ag_ctx, l0.attn.rs_ctx
            l.mlp._init_ctx(max_M)

    def init_triton_dist_AR_ctx(self, max_M: int = 128, ar_method=None):
        from ..ops import comm
        l0 = self.layers[0]
        l0.attn._init_AR_ctx(max_M, ar_method or comm.AllReduceMethod.Unknown, self.dtype)
        for l in self.layers:
            l.attn.ar_ctx, l.attn.ar_method = l0.attn.ar_ctx, l0.attn.ar_method
            l.mlp._init_ctx(max_M)

    init_triton_dist_gemm_ar_ctx = init_triton_dist_AR_ctx

    def finalize(self):
        for l in self.layers:
            l.mlp.finalize()
        l0 = self.layers[0]
        for c in (l0.attn.ag_ctx, l0.attn.rs_ctx, l0.attn.ar_ctx, l0.attn.gemm_ar_ctx):
            if c is not None:
                c.finalize()
        for l in self.layers:
            l.attn.ag_ctx = l.attn.rs_ctx = l.attn.ar_ctx = l.attn.gemm_ar_ctx = None
