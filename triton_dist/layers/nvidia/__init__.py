"""``triton_dist.layers.nvidia`` -- the reference's layer namespace, re-exported from :mod:`triton_dist.parallel`."""
from ...parallel.tp_attn import TP_Attn  # noqa: F401
from ...parallel.tp_mlp import TP_MLP  # noqa: F401
from ...parallel.tp_moe import TP_MoE  # noqa: F401
from ...parallel.ep import (EP_MoE, EPAll2AllLayer, EPConfig, EPLowLatencyAllToAllLayer, EpAll2AllFusedOp,  # noqa: F401
                            TritonDistFusedEpMoeFunction)
from ...parallel.pp import CommOp, PPCommLayer, PyTorchP2P  # noqa: F401
from ...parallel.sp import SpGQAFlashDecodeAttention, UlyssesSPAllToAllLayer  # noqa: F401
from ...parallel.misc import AllGatherLayer, GemmARLayer  # noqa: F401
