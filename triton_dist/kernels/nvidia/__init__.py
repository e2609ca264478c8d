"""``triton_dist.kernels.nvidia`` -- the reference's op namespace (kernels/nvidia/__init__.py:25-101), re-exported
from :mod:`triton_dist.ops` where the sm_100a implementations live."""
from ...ops.ag_gemm import (AllGatherGEMMTensorParallelContext, ag_gemm, create_ag_gemm_context, gemm_non_persistent,  # noqa: F401
                            gemm_persistent)
from ...ops.allgather import (AllGatherMethod, cp_engine_producer_all_gather_inter_node,  # noqa: F401
                              cp_engine_producer_all_gather_intra_node, get_auto_all_gather_method)
from ...ops.comm import (create_fast_allgather_context, fast_allgather, copy_tensor, fill_tensor, reduce_tensor,  # noqa: F401
                         reduce_scatter)
from ...ops.gemm_ar import (create_gemm_ar_context, create_ll_gemm_ar_context, gemm_allreduce_op,  # noqa: F401
                            low_latency_gemm_allreduce_op)
from ...ops.gemm_rs import create_gemm_rs_context, gemm_rs, gemm_rs_mxfp8  # noqa: F401
from ...ops.gemm import gemm as matmul, GemmConfig  # noqa: F401
from ...ops.flash_decode import (gqa_fwd_batch_decode, gqa_fwd_batch_decode_intra_rank,  # noqa: F401
                                 gqa_fwd_batch_decode_persistent)
from ...ops.moe import (ag_group_gemm, create_ag_group_gemm_context, create_moe_ar_context, create_moe_rs_context,  # noqa: F401
                        moe_grouped_gemm, run_moe_reduce_ar, run_moe_reduce_rs)
from ...ops.all_to_all import (all_to_all_post_process, all_to_all_single_2d, all_to_all_vdev_2d, create_all_to_all_context,  # noqa: F401
                               create_all_to_all_single_2d_context, fast_all_to_all)
from ...ops.ep_a2a import combine_kernel_v2, create_ep_ll_a2a_ctx, dispatch_kernel_v2  # noqa: F401
from ...ops.elementwise import swiglu_forward  # noqa: F401
from ...ops.gdn import chunk_gated_delta_rule_fwd  # noqa: F401
from ...parallel.sp import (create_sp_ag_attention_context_intra_node, fused_sp_ag_attn_intra_node)  # noqa: F401
from . import allreduce  # noqa: F401
