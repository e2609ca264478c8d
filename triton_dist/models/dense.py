"""Dense decoder-only LLM (Qwen3 / Llama-3 / Seed-OSS shapes) with tensor-parallel blocks.

Reference: /root/reference/python/triton_dist/models/dense.py:52-235 (``DenseLLMLayer.fwd`` = rmsnorm -> attn ->
residual -> rmsnorm -> mlp -> residual; ``set_fwd`` modes torch | triton_dist | triton_dist_AR |
triton_dist_gemm_ar; one shared ctx per op kind; ``inference`` = embedding -> layers -> norm -> lm_head).
Differences: weights are random-initialised *per shard* straight on the GPU (no network / HF checkpoint here),
the residual add is fused into the next RMSNorm, and every GEMM is the tcgen05 kernel.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import utils as U
from ..ops import comm
from ..ops.elementwise import rmsnorm
from ..parallel.tp_attn import TP_Attn
from ..parallel.tp_mlp import TP_MLP, _linear
from .config import ArchConfig, ModelConfig
from .kv_cache import KV_Cache

FWD_MODES = ("torch", "triton_dist", "triton_dist_AR", "triton_dist_gemm_ar")


def _rand(shape, dtype, device, gen, std=0.02):
    return (torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * std).to(dtype)


class DenseLLMLayer:
    def __init__(self, layer_idx: int, group, rank: int, world_size: int):
        self.layer_idx = layer_idx
        self.attn = TP_Attn(rank, world_size, group)
        self.mlp = TP_MLP(rank, world_size, group)
        self.input_norm_w = self.post_norm_w = None
        self.eps = 1e-6
        self.mode = "torch"

    def init_random(self, arch: ArchConfig, dtype, device, seed: int, rank: int, world: int):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + self.layer_idx * 101 + rank)      # per-rank shard
        H, D = arch.hidden_size, arch.head_dim
        hq, hkv = arch.num_attention_heads // world, max(1, arch.num_key_value_heads // world)
        wqkv = _rand(((hq + 2 * hkv) * D, H), dtype, device, g)
        wo = _rand((H, hq * D), dtype, device, g)
        gate_up = _rand((2 * arch.intermediate_size // world, H), dtype, device, g)
        down = _rand((H, arch.intermediate_size // world), dtype, device, g)
        gr = torch.Generator(device=device)
        gr.manual_seed(seed * 7919 + self.layer_idx)                       # replicated tensors: same on all ranks
        ones = lambda n: (1.0 + 0.02 * torch.randn(n, generator=gr, device=device)).to(dtype)
        qn, kn = (ones(D), ones(D)) if arch.qk_norm else (None, None)
        self.attn._init_parameters_from_shards(wqkv, wo, qn, kn, arch.num_attention_heads, arch.num_key_value_heads, D,
                                               arch.rms_norm_eps, arch.rope_theta)
        self.mlp._init_parameters_from_shards(gate_up, down)
        self.input_norm_w, self.post_norm_w = ones(H), ones(H)
        self.eps = arch.rms_norm_eps

    def init_from_hf(self, hf_layer, arch: ArchConfig, device):
        self.attn._init_parameters(hf_layer.self_attn)
        self.mlp._init_parameters(hf_layer.mlp)
        self.input_norm_w = hf_layer.input_layernorm.weight.detach().to(device)
        self.post_norm_w = hf_layer.post_attention_layernorm.weight.detach().to(device)
        self.eps = arch.rms_norm_eps

    def set_fwd(self, mode: str):
        assert mode in FWD_MODES, mode
        self.mode = mode

    @torch.inference_mode()
    def fwd(self, hidden: torch.Tensor, residual: Optional[torch.Tensor], position_ids, kv_cache: KV_Cache):
        """Returns (delta, residual): the caller (next norm) adds them -- residual add is fused into RMSNorm."""
        if residual is None:
            x, residual = rmsnorm(hidden, self.input_norm_w, self.eps), hidden
        else:
            x, residual = rmsnorm(hidden, self.input_norm_w, self.eps, residual=residual)
        attn_fn = {"torch": self.attn.torch_fwd, "triton_dist": self.attn.dist_triton_fwd,
                   "triton_dist_AR": self.attn.dist_triton_AR_fwd, "triton_dist_gemm_ar": self.attn.dist_triton_gemm_ar_fwd}[self.mode]
        mlp_fn = {"torch": self.mlp.torch_fwd, "triton_dist": self.mlp.dist_triton_fwd,
                  "triton_dist_AR": self.mlp.dist_triton_AR_fwd, "triton_dist_gemm_ar": self.mlp.dist_triton_gemm_ar_fwd}[self.mode]
        a = attn_fn(x, position_ids, kv_cache, self.layer_idx)
        x, residual = rmsnorm(a, self.post_norm_w, self.eps, residual=residual)
        m = mlp_fn(x)
        return m, residual


class DenseLLM:
    def __init__(self, model_config: ModelConfig, group=None):
        self.model_config = model_config
        self.arch = model_config.arch()
        self.rank, self.world_size = model_config.rank, model_config.world_size
        self.group = group
        self.dtype = model_config.dtype
        self.device = U.current_device()
        layer_cls = getattr(self, "_layer_cls", DenseLLMLayer)
        self.layers = [layer_cls(i, group, self.rank, self.world_size) for i in range(self.arch.num_hidden_layers)]
        self.embed_tokens = self.lm_head = self.norm_w = None
        self.mode = "torch"
        self.num_layers = self.arch.num_hidden_layers
        self.num_key_value_heads = self.arch.num_key_value_heads
        self.head_dim = self.arch.head_dim
        self.max_length = model_config.max_length
        self.init_parameters()

    # ---- weights ------------------------------------------------------------------------------------------
    def init_parameters(self):
        a, mc = self.arch, self.model_config
        if not mc.random_init:
            return self._init_from_hf()
        g = torch.Generator(device=self.device)
        g.manual_seed(mc.seed)
        self.embed_tokens = _rand((a.vocab_size, a.hidden_size), self.dtype, self.device, g)
        self.lm_head = self.embed_tokens if a.tie_word_embeddings else _rand((a.vocab_size, a.hidden_size), self.dtype, self.device, g)
        self.norm_w = torch.ones(a.hidden_size, dtype=self.dtype, device=self.device)
        for layer in self.layers:
            layer.init_random(a, self.dtype, self.device, mc.seed, self.rank, self.world_size)

    def _init_from_hf(self):
        from transformers import AutoModelForCausalLM
        hf = AutoModelForCausalLM.from_pretrained(self.model_config.model_name, torch_dtype=self.dtype,
                                                  local_files_only=self.model_config.local_only)
        self.embed_tokens = hf.model.embed_tokens.weight.detach().to(self.device)
        self.lm_head = hf.lm_head.weight.detach().to(self.device)
        self.norm_w = hf.model.norm.weight.detach().to(self.device)
        for layer, hl in zip(self.layers, hf.model.layers):
            layer.init_from_hf(hl, self.arch, self.device)
        del hf

    # ---- modes / contexts ---------------------------------------------------------------------------------
    def set_fwd(self, mode: str = "torch"):
        self.mode = mode
        for l in self.layers:
            l.set_fwd(mode)

    def init_triton_dist_ctx(self, max_M: int = 4096):
        """One AG/RS context per op kind, created on layer 0 and shared by all layers (dense.py:169-188): the
        contexts are phase-counted and double buffered, so back-to-back layers can reuse them safely."""
        l0 = self.layers[0]
        l0.attn._init_ctx(max_M)
        l0.mlp._init_ctx(max_M)
        for l in self.layers[1:]:
            l.attn.ag_ctx, l.attn.rs_ctx = l0.attn.ag_ctx, l0.attn.rs_ctx
            l.mlp.ag_ctx, l.mlp.rs_ctx = l0.mlp.ag_ctx, l0.mlp.rs_ctx

    def init_triton_dist_AR_ctx(self, max_M: int = 128, ar_method=comm.AllReduceMethod.Unknown):
        l0 = self.layers[0]
        l0.attn._init_AR_ctx(max_M, ar_method, self.dtype)
        for l in self.layers:
            l.attn.ar_ctx = l.mlp.ar_ctx = l0.attn.ar_ctx
            l.attn.ar_method = l.mlp.ar_method = ar_method

    def init_triton_dist_gemm_ar_ctx(self, max_M: int = 128):
        l0 = self.layers[0]
        l0.attn._init_gemm_ar_ctx(max_M, self.dtype)
        for l in self.layers:
            l.attn.gemm_ar_ctx = l.mlp.gemm_ar_ctx = l0.attn.gemm_ar_ctx

    def finalize(self):
        l0 = self.layers[0]
        seen = set()
        for c in (l0.attn.ag_ctx, l0.attn.rs_ctx, l0.mlp.ag_ctx, l0.mlp.rs_ctx, l0.attn.ar_ctx, l0.attn.gemm_ar_ctx):
            if c is not None and id(c) not in seen:
                seen.add(id(c))
                c.finalize()
        for l in self.layers:
            l.attn.ag_ctx = l.attn.rs_ctx = l.attn.ar_ctx = l.attn.gemm_ar_ctx = None
            l.mlp.ag_ctx = l.mlp.rs_ctx = l.mlp.ar_ctx = l.mlp.gemm_ar_ctx = None

    # ---- forward ------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def inference(self, input_ids: torch.Tensor, position_ids: torch.Tensor, kv_cache: KV_Cache) -> torch.Tensor:
        """input_ids/position_ids: [bsz, q_len] ([bsz/W, q_len] token rows in ``triton_dist`` mode -- positions stay
        full ``[bsz, q_len]`` because attention runs on the gathered batch).  Returns fp32 logits of the last token."""
        h = torch.nn.functional.embedding(input_ids, self.embed_tokens)
        delta, residual = h, None
        for layer in self.layers:
            delta, residual = layer.fwd(delta, residual, position_ids, kv_cache)
        h, _ = rmsnorm(delta, self.norm_w, self.arch.rms_norm_eps, residual=residual)
        if h.shape[1] > 1:
            h = h[:, -1:, :]
        logits = _linear(h.reshape(-1, h.shape[-1]).contiguous(), self.lm_head)
        return logits.float()
