"""``triton_dist.kernels.nvidia.allreduce`` (reference file of the same name: create_allreduce_ctx :109,
all_reduce :1130, get_auto_allreduce_method :1102)."""
from ...ops.comm import (AllReduceContext, all_reduce, create_allreduce_ctx, get_auto_allreduce_method)  # noqa: F401
