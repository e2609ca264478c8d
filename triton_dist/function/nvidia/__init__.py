"""Autograd functions (reference: python/triton_dist/function/nvidia/)."""
from ...parallel.ep import TritonDistFusedEpMoeFunction  # noqa: F401  (low-latency exchange, inference-grade forward)
from .ep_moe_fused import MegaEpMoeFunction, mega_ep_moe_autograd  # noqa: F401  (Mega-EP kernels, full forward + backward)
