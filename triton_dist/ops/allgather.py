"""Copy-engine all-gather producers (SM-free transport) and method selection.

Reference: /root/reference/python/triton_dist/kernels/nvidia/allgather.py:46-124,202 -- full-mesh pull/push and
ring variants driven from the host with ``cudaMemcpyAsync`` + ``cuStreamWriteValue`` flags.  On an NVSwitch box
full-mesh is the fast choice, so ``get_auto_all_gather_method`` answers All2All_IntraNode; the 1-D and 2-D ring
producers are implemented as well (same contract: segment s of my buffer is valid once ``flag[s] == signal_value``):
they move every byte W-1 (1-D) or G-1 + 1 (2-D) hops and exist for links where a ring is the topology.
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


def cp_engine_producer_all_gather_ring_push_1d(rank: int, num_ranks: int, local_tensor: torch.Tensor,
                                               remote_tensor_buffers: List[torch.Tensor], barrier_buffers: List[torch.Tensor],
                                               stream=None, signal_value: int = 1):
    """1-D push ring (reference allgather.py ring_push_1d): at step j every rank forwards segment ``(rank - j) % W`` -- its own
    shard first, then what it received in the previous step -- into the NEXT rank's buffer (copy engine, peer mapping) and
    release-sets ``flag[segment]`` there; a segment is forwarded only after its own arrival flag has been acquired on the stream."""
    from .. import language as dl
    W, M_per = num_ranks, local_tensor.shape[0]
    nxt = (rank + 1) % W
    mine, dst = remote_tensor_buffers[rank], remote_tensor_buffers[nxt]
    cm = torch.cuda.stream(stream) if (local_tensor.is_cuda and stream is not None) else None
    if cm is not None:
        cm.__enter__()
    try:
        seg = slice(rank * M_per, (rank + 1) * M_per)
        mine[seg].copy_(local_tensor)
        U.set_signal(barrier_buffers[rank][rank:rank + 1], signal_value, stream)
        for j in range(W - 1):
            s = (rank - j) % W
            sl = slice(s * M_per, (s + 1) * M_per)
            if j > 0:
                dl.wait(barrier_buffers[rank][s:s + 1], 1, wait_value=signal_value, geq=True)        # arrived from rank-1
            dst[sl].copy_(mine[sl], non_blocking=True)
            dl.notify(barrier_buffers[rank][s:s + 1], nxt, signal=signal_value, sig_op="set")
        last = (rank + 1) % W                                  # the segment that reaches me in the final step
        if W > 1:
            dl.wait(barrier_buffers[rank][last:last + 1], 1, wait_value=signal_value, geq=True)
    finally:
        if cm is not None:
            cm.__exit__(None, None, None)
    return mine


def cp_engine_producer_all_gather_ring_push_2d(rank: int, num_ranks: int, local_tensor: torch.Tensor,
                                               remote_tensor_buffers: List[torch.Tensor], barrier_buffers: List[torch.Tensor],
                                               stream=None, signal_value: int = 1, group_size: Optional[int] = None):
    """2-D ring (reference ring_push_2d: intra-NUMA ring, then across): ranks form groups of ``group_size`` (default W / 2); every
    segment first travels the ring of its source's group, and each rank forwards every segment of its group -- its own and the
    ones it receives -- to its partner ranks (same index) in the other groups.  G-1 + 1 hops instead of W-1."""
    from .. import language as dl
    W, M_per = num_ranks, local_tensor.shape[0]
    G = group_size or (W // 2 if W % 2 == 0 and W >= 4 else W)
    if G >= W or W % G:
        return cp_engine_producer_all_gather_ring_push_1d(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, stream,
                                                          signal_value)
    grp, idx = rank // G, rank % G
    nxt = grp * G + (idx + 1) % G
    partners = [g * G + idx for g in range(W // G) if g != grp]
    mine = remote_tensor_buffers[rank]
    cm = torch.cuda.stream(stream) if (local_tensor.is_cuda and stream is not None) else None
    if cm is not None:
        cm.__enter__()
    try:
        mine[rank * M_per:(rank + 1) * M_per].copy_(local_tensor)
        U.set_signal(barrier_buffers[rank][rank:rank + 1], signal_value, stream)
        for j in range(G):
            s = grp * G + (idx - j) % G                        # segment of my group that is in my buffer at step j
            sl = slice(s * M_per, (s + 1) * M_per)
            if j > 0:
                dl.wait(barrier_buffers[rank][s:s + 1], 1, wait_value=signal_value, geq=True)
            if j < G - 1:                                      # around my group's ring
                remote_tensor_buffers[nxt][sl].copy_(mine[sl], non_blocking=True)
                dl.notify(barrier_buffers[rank][s:s + 1], nxt, signal=signal_value, sig_op="set")
            for p_ in partners:                                # and across to the same index of the other groups
                remote_tensor_buffers[p_][sl].copy_(mine[sl], non_blocking=True)
                dl.notify(barrier_buffers[rank][s:s + 1], p_, signal=signal_value, sig_op="set")
        for s in range(W):                                     # everything that is sent to me has arrived
            if s != rank:
                dl.wait(barrier_buffers[rank][s:s + 1], 1, wait_value=signal_value, geq=True)
    finally:
        if cm is not None:
            cm.__exit__(None, None, None)
    return mine


def cp_engine_producer_all_gather(rank: int, num_ranks: int, local_tensor: torch.Tensor, remote_tensor_buffers, barrier_buffers,
                                  stream=None, signal_value: int = 1, method: AllGatherMethod = AllGatherMethod.Auto):
    """Dispatch on :class:`AllGatherMethod` (reference allgather.py:202)."""
    if method in (AllGatherMethod.Auto, AllGatherMethod.All2All_IntraNode):
        return cp_engine_producer_all_gather_intra_node(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, stream, signal_value)
    if method == AllGatherMethod.Ring1D_IntraNode:
        return cp_engine_producer_all_gather_ring_push_1d(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, stream, signal_value)
    if method == AllGatherMethod.Ring2D_IntraNode:
        return cp_engine_producer_all_gather_ring_push_2d(rank, num_ranks, local_tensor, remote_tensor_buffers, barrier_buffers, stream, signal_value)
    raise NotImplementedError(f"{method}: inter-node transports are out of scope for a single NVSwitch domain (SURVEY.md N9)")


def cp_engine_producer_all_gather_inter_node(*_a, **_k):
    raise NotImplementedError("inter-node transports are out of scope for a single NVSwitch domain (SURVEY.md N9)")
