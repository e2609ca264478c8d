"""Pipeline-parallel stage hand-off (reference: layers/nvidia/p2p.py:CommOp, layers/nvidia/pp_block.py:PPCommLayer).

``CommOp``: read / write / set_signal / wait_signal on symmetric buffers.  ``PPCommLayer.send / recv`` moves an
activation tensor to the next stage either through the symmetric heap (``backend="triton_dist"``: one peer-mapped
copy + one release flag, no NCCL rendezvous) or with NCCL send/recv (``backend="torch"``, the baseline)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import utils as U
from ..ops import p2p


class CommOp:
    def __init__(self, max_numel: int, dtype: torch.dtype, num_buffers: int = 2):
        self.num_buffers = num_buffers
        self.buffers = U.nvshmem_create_tensor((num_buffers, max_numel), dtype)
        self.signals = U.nvshmem_create_tensor((max(num_buffers, 8),), torch.int32)
        self.acks = U.nvshmem_create_tensor((max(num_buffers, 8),), torch.int32)
        U.barrier_all_host()

    def write(self, peer: int, slot: int, src: torch.Tensor):
        p2p.p2p_put(self.buffers[slot], src.reshape(-1), peer)

    def read(self, slot: int, numel: int, peer: Optional[int] = None) -> torch.Tensor:
        if peer is None:
            return self.buffers[slot, :numel]
        out = torch.empty(numel, dtype=self.buffers.dtype, device=self.buffers.device)
        p2p.p2p_get(out, self.buffers[slot], peer)
        return out

    def set_signal(self, peer: int, slot: int, value: int):
        p2p.p2p_set_signal(self.signals[slot:slot + 1], peer, value)

    def wait_signal(self, slot: int, value: int):
        p2p.p2p_wait_signal(self.signals[slot:slot + 1], value)

    def finalize(self):
        for t in (self.buffers, self.signals, self.acks):
            U.nvshmem_free_tensor_sync(t)


class PPCommLayer:
    """send/recv between pipeline stages.  Slots are used round-robin; the receiver acknowledges a slot so the sender
    can reuse it (credit flow control), all with monotonically increasing sequence numbers (no resets)."""

    def __init__(self, max_numel: int, dtype: torch.dtype, rank: int, world_size: int, backend: str = "triton_dist",
                 num_buffers: int = 2, group=None):
        self.rank, self.world_size, self.backend, self.group = rank, world_size, backend, group
        self.op = CommOp(max_numel, dtype, num_buffers) if backend == "triton_dist" else None
        self.num_buffers = num_buffers
        self.send_seq = 0
        self.recv_seq = 0

    def send(self, x: torch.Tensor, dst: int):
        if self.backend == "torch":
            dist.send(x.contiguous(), dst=dist.get_global_rank(self.group, dst) if self.group is not None else dst, group=self.group)
            return
        self.send_seq += 1
        slot = self.send_seq % self.num_buffers
        if self.send_seq > self.num_buffers:       # wait until the receiver has consumed the previous use of this slot
            p2p.p2p_wait_signal(self.op.acks[slot:slot + 1], self.send_seq - self.num_buffers)
        self.op.write(dst, slot, x)
        self.op.set_signal(dst, slot, self.send_seq)

    def recv(self, shape, dtype, src: int) -> torch.Tensor:
        if self.backend == "torch":
            out = torch.empty(shape, dtype=dtype, device=U.current_device())
            dist.recv(out, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
            return out
        self.recv_seq += 1
        slot = self.recv_seq % self.num_buffers
        self.op.wait_signal(slot, self.recv_seq)
        numel = 1
        for s in shape:
            numel *= s
        out = self.op.read(slot, numel).view(shape).clone()
        p2p.p2p_set_signal(self.op.acks[slot:slot + 1], src, self.recv_seq)
        return out

    def finalize(self):
        if self.op is not None:
            self.op.finalize()


PyTorchP2P = PPCommLayer
