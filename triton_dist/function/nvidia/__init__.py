"""Autograd functions (reference: python/triton_dist/function/nvidia/)."""
from ...parallel.ep import TritonDistFusedEpMoeFunction  # noqa: F401  (low-latency exchange, inference-grade forward)
from .ep_moe_fused import MegaEpMoeFunction, mega_ep_moe_autograd  # noqa: F401  (Mega-EP kernels, full forward + backward)
from .common import (MoEOptimConfig, TritonDistEpContext, custom_bwd, custom_fwd, deinit_triton_dist_ep_op, fused_ep_moe,  # noqa: F401
                     get_ep_capacity, get_moe_optim_config, get_triton_dist_ep_op, get_triton_dist_ep_stream,
                     get_triton_dist_moe_profile_enabled, get_triton_dist_profile_output_dir, init_triton_dist_ep_ctx,
                     init_triton_dist_ep_op, set_triton_dist_moe_profile_enabled, triton_dist_ep_op_initialized)
