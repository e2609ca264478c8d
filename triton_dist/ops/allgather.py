"""Copy-engine all-gather producers (SM-free transport) and method selection.

Reference: /root/reference/python/triton_dist/kernels/nvidia/allgather.py:46-124,202 -- full-mesh pull/push and
ring variants driven from the host with ``cudaMemcpyAsync`` + ``cuStreamWriteValue`` flags.  On an NVSwitch box
every method except full-mesh is pointless, so ``get_auto_all_gather_method`` always answers All2All_IntraNode;
the ring enums are kept for API parity and map to the same implementation.
"""
from __future__ import annotations

import enum
from typing import List, Optional, Tuple

import torch

from .. import utils as U


class AllGatherMethod(enum.Enum):
    Auto = 0
    All2All_IntraNode = 1
    All2All_InterNode = 2
    Ring1D_IntraNode = 3
    Ring2D_IntraNode = 4
    Ring1D_InterNode = 5
    Ring2D_InterNode = 6


def get_auto_all_gather_method(num_ranks: int, num_local_ranks: int) -> AllGatherMethod:
    if num_ranks != num_local_ranks:
        raise NotImplementedError("inter-node all-gather is out of scope (single NVSwitch domain)")
    return AllGatherMethod.All2All_IntraNode


def create_allgather_buffers(M: int, N: int, dtype: torch.dtype) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """Symmetric ``[M, N]`` data buffers + ``[world]`` int32 flag arrays, as lists of peer views (index = rank)."""
    W = U.world_size()
    bufs = U.nvshmem_create_tensors((M, N), dtype, U.rank(), W)
    flags = U.nvshmem_create_tensors((max(W, 8),), torch.int32, U.rank(), W)
    U.barrier_all_host()
    return bufs, flags


def cp_engine_producer_all_gather_intra_node(rank: int, num_ranks: int, local_tensor: torch.Tensor,
                                             remote_tensor_buffers: List[torch.Tensor], barrier_buffers: List[torch.Tensor],
                                             stream=None, signal_value: int = 1,
                                             method: AllGatherMethod = AllGatherMethod.All2All_IntraNode):
    """Full-mesh *pull*: my shard goes into my buffer, then for every other source I copy its segment out of ITS
    buffer and set ``flag[src]`` on my side once the segment is resident (stream-ordered flag write)."""
    M_per = local_tensor.shape[0]
    mine = remote_tensor_buffers[rank]
    is_cuda = local_tensor.is_cuda
    ctx_stream = torch.cuda.stream(stream) if (is_cuda and stream is not None) else None
    if ctx_stream is not None:
        ctx_stream.__enter__()
    try:
        mine[rank * M_per:(rank + 1) * M_per].copy_(local_tensor)
        U.set_signal(barrier_buffers[rank][rank:rank + 1], signal_value, stream)
        # peers must have staged their shard before I read it
        U.barrier_all_on_stream(stream=stream)
        for j in range(1, num_ranks):
            src = (rank + j) % num_ranks
            seg = slice(src * M_per, (src + 1) * M_per)
            mine[seg].copy_(remote_tensor_buffers[src][seg], non_blocking=True)
            U.set_signal(barrier_buffers[rank][src:src + 1], signal_value, stream)
    finally:
        if ctx_stream is not None:
            ctx_stream.__exit__(None, None, None)
    return mine


def cp_engine_producer_all_gather_inter_node(*_a, **_k):
    raise NotImplementedError("inter-node transports are out of scope for a single NVSwitch domain (SURVEY.md N9)")
