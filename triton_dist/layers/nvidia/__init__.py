"""``triton_dist.layers.nvidia`` -- the reference's layer namespace, re-exported from :mod:`triton_dist.parallel`."""
from ...parallel.tp_attn import TP_Attn  # noqa: F401
from ...parallel.tp_mlp import TP_MLP  # noqa: F401
