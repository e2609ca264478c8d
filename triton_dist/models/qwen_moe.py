"""Qwen3-MoE (TP-MoE blocks).  Filled in with the MoE ops (see triton_dist/parallel/tp_moe.py)."""
from .dense import DenseLLM


class Qwen3MoE(DenseLLM):
    pass
