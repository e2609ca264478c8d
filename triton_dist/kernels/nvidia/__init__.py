"""``triton_dist.kernels.nvidia`` -- the reference's op namespace (kernels/nvidia/__init__.py:25-101), re-exported
from :mod:`triton_dist.ops` where the sm_100a implementations live."""
from ...ops.ag_gemm import (AllGatherGEMMTensorParallelContext, ag_gemm, create_ag_gemm_context, gemm_non_persistent,  # noqa: F401
                            gemm_persistent)
from ...ops.allgather import (AllGatherMethod, cp_engine_producer_all_gather_inter_node,  # noqa: F401
                              cp_engine_producer_all_gather_intra_node, get_auto_all_gather_method)
from ...ops.comm import (create_fast_allgather_context, fast_allgather)  # noqa: F401
from ...ops.gemm_ar import (create_gemm_ar_context, create_ll_gemm_ar_context, gemm_allreduce_op,  # noqa: F401
                            low_latency_gemm_allreduce_op)
from ...ops.gemm_rs import create_gemm_rs_context, gemm_rs  # noqa: F401
from ...ops.flash_decode import (gqa_fwd_batch_decode, gqa_fwd_batch_decode_intra_rank,  # noqa: F401
                                 gqa_fwd_batch_decode_persistent)
from . import allreduce  # noqa: F401
