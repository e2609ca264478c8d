"""Model configuration (reference: /root/reference/python/triton_dist/models/config.py:30-37) plus a table of the
public architectures the reference's AutoLLM maps (models/__init__.py:33-69), so the random-weight demo needs no
network / HF download (the reference reads the same numbers through ``AutoConfig``, dense.py:126-135)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch


@dataclass
class ArchConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    vocab_size: int
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    qk_norm: bool = True
    tie_word_embeddings: bool = False
    # MoE (0 experts = dense)
    num_experts: int = 0
    num_experts_per_tok: int = 0
    moe_intermediate_size: int = 0
    norm_topk_prob: bool = True


ARCHS = {
    "Qwen/Qwen3-0.6B": ArchConfig(1024, 3072, 28, 16, 8, 128, 151936, tie_word_embeddings=True),
    "Qwen/Qwen3-8B": ArchConfig(4096, 12288, 36, 32, 8, 128, 151936),
    "Qwen/Qwen3-14B": ArchConfig(5120, 17408, 40, 40, 8, 128, 151936),
    "Qwen/Qwen3-32B": ArchConfig(5120, 25600, 64, 64, 8, 128, 151936),
    "meta-llama/Meta-Llama-3-70B": ArchConfig(8192, 28672, 80, 64, 8, 128, 128256, 1e-5, 5e5, qk_norm=False),
    "ByteDance-Seed/Seed-OSS-36B-Instruct": ArchConfig(5120, 27648, 64, 80, 8, 128, 155136, 1e-6, 1e7, qk_norm=False),
    "Qwen/Qwen3-30B-A3B": ArchConfig(2048, 6144, 48, 32, 4, 128, 151936, num_experts=128, num_experts_per_tok=8,
                                     moe_intermediate_size=768),
    "Qwen/Qwen3-235B-A22B": ArchConfig(4096, 12288, 94, 64, 4, 128, 151936, num_experts=128, num_experts_per_tok=8,
                                       moe_intermediate_size=1536),
    # tiny configs for tests / smoke
    "tiny-dense": ArchConfig(256, 512, 2, 8, 8, 128, 1024),
    "tiny-moe": ArchConfig(256, 512, 2, 8, 8, 128, 1024, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=256),
}


def arch_from_hf_config(path: str) -> ArchConfig:
    """ArchConfig of a local HF checkpoint directory (``config.json``), the way the reference reads ``AutoConfig``
    (models/dense.py:126-135)."""
    import json
    import os
    c = json.load(open(os.path.join(path, "config.json")))
    heads = c["num_attention_heads"]
    return ArchConfig(
        hidden_size=c["hidden_size"], intermediate_size=c.get("intermediate_size", 0), num_hidden_layers=c["num_hidden_layers"],
        num_attention_heads=heads, num_key_value_heads=c.get("num_key_value_heads", heads),
        head_dim=c.get("head_dim") or c["hidden_size"] // heads, vocab_size=c["vocab_size"],
        rms_norm_eps=c.get("rms_norm_eps", 1e-6), rope_theta=c.get("rope_theta", 1e6),
        qk_norm="qwen3" in c.get("model_type", "").lower(), tie_word_embeddings=c.get("tie_word_embeddings", False),
        num_experts=c.get("num_experts", 0) or 0, num_experts_per_tok=c.get("num_experts_per_tok", 0) or 0,
        moe_intermediate_size=c.get("moe_intermediate_size", 0) or 0, norm_topk_prob=c.get("norm_topk_prob", True))


@dataclass
class ModelConfig:
    model_name: str = "Qwen/Qwen3-32B"
    max_length: int = 4096
    dtype: torch.dtype = torch.bfloat16
    local_only: bool = True
    rank: int = 0
    world_size: int = 1
    random_init: bool = True          # no network here: weights are seeded random tensors of the right shapes
    seed: int = 1234
    num_layers_override: Optional[int] = None

    def arch(self) -> ArchConfig:
        import os
        if self.model_name not in ARCHS and os.path.isfile(os.path.join(self.model_name, "config.json")):
            a = arch_from_hf_config(self.model_name)          # a local HF checkpoint directory (weights are loaded from it)
        elif self.model_name not in ARCHS:
            raise KeyError(f"unknown architecture '{self.model_name}'; known: {sorted(ARCHS)} (or a local HF checkpoint directory)")
        else:
            a = ARCHS[self.model_name]
        if self.num_layers_override:
            import dataclasses
            a = dataclasses.replace(a, num_hidden_layers=self.num_layers_override)
        return a
