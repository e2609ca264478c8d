"""``triton_dist.kernels.allreduce`` (reference: kernels/allreduce.py:31-80)."""
from ..ops.comm import (AllReduceMethod, OverlappingAllReduceMethod, get_allreduce_methods,  # noqa: F401
                        get_auto_all_reduce_method, to_allreduce_method)
