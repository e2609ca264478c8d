"""Parallel layers: TP (AG-GEMM/GEMM-RS, GEMM-AR), TP-MoE, EP, SP, PP."""
from .tp_attn import TP_Attn  # noqa: F401
from .tp_mlp import TP_MLP  # noqa: F401
