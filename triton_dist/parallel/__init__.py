"""Parallel layers: TP (AG-GEMM/GEMM-RS, GEMM-AR), TP-MoE, EP, SP, PP."""
from .tp_attn import TP_Attn  # noqa: F401
from .tp_mlp import TP_MLP  # noqa: F401
from .tp_moe import TP_MoE  # noqa: F401,E402
from .ep import (EP_MoE, EPAll2AllLayer, EPConfig, EPLowLatencyAllToAllLayer, EPNormalAll2AllLayer, EpAll2AllFusedOp,
                 TritonDistFusedEpMoeFunction)  # noqa: F401,E402
from .pp import CommOp, PPCommLayer  # noqa: F401,E402
from .sp import SpGQAFlashDecodeAttention, UlyssesSPAllToAllLayer, fused_sp_ag_attn_intra_node  # noqa: F401,E402
from .misc import AllGatherLayer, GemmARLayer  # noqa: F401,E402
