"""All-to-all family over the symmetric heap (one push kernel, device-side splits, no barrier).

Reference:
  * ``create_all_to_all_context`` / ``fast_all_to_all`` / ``all_to_all_post_process``  (EP, DeepEP-LL-like)
    -- kernels/nvidia/low_latency_all_to_all.py:33-279
  * ``all_to_all_single_2d`` (+ctx) -- kernels/nvidia/all_to_all_single_2d.py:43-205
  * ``all_to_all_vdev_2d`` / ``all_to_all_v_offset_op`` -- kernels/nvidia/all_to_all_vdev_2d_offset.py
Kernel: csrc/comm_kernels.cu ``all_to_all_kernel``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from .. import _C
from .. import utils as U
from .comm import SymmArgs, symm_args

c_void_p, c_ll, c_int = C.c_void_p, C.c_longlong, C.c_int


class _A2AArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("g", c_ll), ("row_bytes", c_ll), ("max_rows", c_ll), ("row_bytes2", c_ll), ("grid", c_ll),
                ("send", c_void_p), ("cum", c_void_p), ("send2", c_void_p), ("recv_buf", c_void_p), ("recv_buf_bytes", c_ll),
                ("recv_buf2", c_void_p), ("recv_buf2_bytes", c_ll), ("recv_meta", c_void_p), ("flags", c_void_p), ("phase", c_void_p)]


_C.register("td_all_to_all", c_int, [C.POINTER(_A2AArgs), c_void_p])


@dataclass
class AllToAllContext:
    max_m: int                  # max rows one rank sends to one destination
    hidden: int
    rank: int
    num_tot_experts: int
    world_size: int
    experts_per_rank: int
    dtype: torch.dtype
    scale_dtype: Optional[torch.dtype]
    recv_buf: torch.Tensor = None      # symmetric [2, W, max_m, hidden]
    recv_scale: torch.Tensor = None    # symmetric [2, W, max_m, scale_cols] or None
    recv_meta: torch.Tensor = None     # symmetric int32 [2, W, g + 1]
    flags: torch.Tensor = None         # symmetric int32 [2, W]
    phase: torch.Tensor = None
    scale_cols: int = 0
    host_calls: int = 0

    def finalize(self):
        heap = U.get_heap()
        for t in (self.recv_buf, self.recv_scale, self.recv_meta, self.flags):
            if t is not None:
                heap.free_tensor(t)
        self.recv_buf = self.recv_scale = self.recv_meta = self.flags = None


def create_all_to_all_context(max_m: int, hidden: int, rank: Optional[int] = None, num_tot_experts: Optional[int] = None,
                              world_size: Optional[int] = None, experts_per_rank: Optional[int] = None,
                              dtype: torch.dtype = torch.bfloat16, scale_dtype: Optional[torch.dtype] = None,
                              scale_cols: int = 0) -> AllToAllContext:
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    world_size = heap.world if world_size is None else world_size
    experts_per_rank = experts_per_rank or ((num_tot_experts or world_size) // world_size)
    num_tot_experts = num_tot_experts or experts_per_rank * world_size
    ctx = AllToAllContext(max_m, hidden, rank, num_tot_experts, world_size, experts_per_rank, dtype, scale_dtype, scale_cols=scale_cols)
    ctx.recv_buf = heap.tensor((2, world_size, max_m, hidden), dtype)
    if scale_dtype is not None and scale_cols > 0:
        ctx.recv_scale = heap.tensor((2, world_size, max_m, scale_cols), scale_dtype)
    ctx.recv_meta = heap.tensor((2, world_size, experts_per_rank + 1), torch.int32)
    ctx.flags = heap.tensor((2, max(world_size, 4)), torch.int32)
    ctx.phase = torch.zeros(4, dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


def fast_all_to_all(ctx: AllToAllContext, send_tensor: torch.Tensor, send_split_cumsum: torch.Tensor,
                    send_scale: Optional[torch.Tensor] = None, num_sms: int = 32):
    """send_tensor: [rows, hidden] sorted by destination expert; send_split_cumsum: int32 [E + 1] (device);
    Returns (recv_splits [W, experts_per_rank] int32, recv_tensor [W, max_m, hidden], recv_scale | None) -- views of the
    symmetric receive buffer of this call's parity (valid until the call after next).  While a CUDA graph is being captured
    the results are device-selected COPIES of the half the kernel wrote (replay safe)."""
    W, g = ctx.world_size, ctx.experts_per_rank
    assert send_tensor.is_contiguous() and send_split_cumsum.numel() == W * g + 1
    cum = send_split_cumsum.to(torch.int32).contiguous()
    if not send_tensor.is_cuda:
        return _a2a_host(ctx, send_tensor, cum, send_scale)
    ph = ctx.host_calls + 1
    par = ph & 1
    a = _A2AArgs()
    a.symm = symm_args()
    a.g, a.row_bytes, a.max_rows = g, send_tensor.shape[1] * send_tensor.element_size(), ctx.max_m
    a.grid = num_sms
    a.send, a.cum = send_tensor.data_ptr(), cum.data_ptr()
    a.recv_buf, a.recv_buf_bytes = ctx.recv_buf.data_ptr(), ctx.recv_buf[0].numel() * ctx.recv_buf.element_size()
    if send_scale is not None:
        assert ctx.recv_scale is not None and send_scale.is_contiguous()
        a.send2, a.row_bytes2 = send_scale.data_ptr(), send_scale.shape[1] * send_scale.element_size()
        assert a.row_bytes2 % 16 == 0, "scale rows must be multiples of 16 bytes"
        a.recv_buf2, a.recv_buf2_bytes = ctx.recv_scale.data_ptr(), ctx.recv_scale[0].numel() * ctx.recv_scale.element_size()
    a.recv_meta, a.flags, a.phase = ctx.recv_meta.data_ptr(), ctx.flags.data_ptr(), ctx.phase.data_ptr()
    _C.check(_C.cuda_lib().td_all_to_all(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_all_to_all")
    ctx.host_calls = ph
    if torch.cuda.is_current_stream_capturing():
        # The kernel picks the receive half from its DEVICE-resident call counter, so a replayed graph alternates halves while
        # a view chosen from the host counter would stay frozen at the capture-time parity.  Under capture the half is
        # therefore selected on the device too: phase[0] (already advanced by the kernel, stream-ordered) & 1.
        sel = (ctx.phase[0:1] & 1).to(torch.int64)
        meta = ctx.recv_meta.index_select(0, sel)[0]
        buf = ctx.recv_buf.index_select(0, sel)[0]
        sc = ctx.recv_scale.index_select(0, sel)[0] if send_scale is not None else None
        return meta[:, :g], buf, sc
    meta = ctx.recv_meta[par]
    return meta[:, :g], ctx.recv_buf[par], (ctx.recv_scale[par] if send_scale is not None else None)


def all_to_all_post_process(ctx: AllToAllContext, recv_splits: torch.Tensor, recv_tensor: torch.Tensor,
                            recv_scale: Optional[torch.Tensor] = None):
    """Compact the padded ``[W, max_m, H]`` receive buffer into contiguous rows (source-major), like the reference."""
    per_src = recv_splits.sum(dim=1)
    mask = torch.arange(ctx.max_m, device=recv_tensor.device)[None, :] < per_src[:, None]
    out = recv_tensor[mask]
    sc = recv_scale[mask] if recv_scale is not None else None
    return (out, sc) if recv_scale is not None else out


def _a2a_host(ctx, send, cum, send_scale):
    import ctypes
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, g = ctx.world_size, ctx.rank, ctx.experts_per_rank
    ctx.host_calls += 1
    ph, par = ctx.host_calls, ctx.host_calls & 1
    for j in range(W):
        dst = (me + j) % W
        r0, r1 = int(cum[dst * g]), int(cum[(dst + 1) * g])
        n = r1 - r0
        heap.peer_view(ctx.recv_buf, dst)[par, me, :n] = send[r0:r1]
        if send_scale is not None:
            heap.peer_view(ctx.recv_scale, dst)[par, me, :n] = send_scale[r0:r1]
        meta = heap.peer_view(ctx.recv_meta, dst)[par, me]
        meta[:g] = (cum[dst * g + 1:(dst + 1) * g + 1] - cum[dst * g:(dst + 1) * g])
        meta[g] = n
        lib.tdh_notify32(ctypes.c_void_p(heap.peer_ptr(ctx.flags[par, me:me + 1].data_ptr(), dst)), ph, 1)
    if lib.tdh_wait32_n(ctypes.c_void_p(ctx.flags[par].data_ptr()), W, ph, 1, 60_000_000):
        raise TimeoutError("all_to_all: a peer never delivered")
    return ctx.recv_meta[par][:, :g], ctx.recv_buf[par], (ctx.recv_scale[par] if send_scale is not None else None)


# ------------------------------------------------------------------------------------------------------------
# fixed-layout all-to-all (Ulysses, all_to_all_single_2d)
# ------------------------------------------------------------------------------------------------------------
def create_all_to_all_single_2d_context(max_rows_per_peer: int, cols: int, dtype: torch.dtype) -> AllToAllContext:
    return create_all_to_all_context(max_rows_per_peer, cols, experts_per_rank=1, dtype=dtype)


def all_to_all_single_2d(ctx: AllToAllContext, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Equal-split all-to-all on dim 0 (``torch.distributed.all_to_all_single`` semantics): x [W * n, cols] -> [W * n, cols]
    where output block s is what rank s sent to me."""
    W = ctx.world_size
    n = x.shape[0] // W
    assert x.shape[0] == n * W and n <= ctx.max_m
    cum = torch.arange(0, (W + 1) * n, n, dtype=torch.int32, device=x.device)
    _, recv, _ = fast_all_to_all(ctx, x.contiguous(), cum)
    res = recv[:, :n].reshape(W * n, -1)
    if out is not None:
        out.copy_(res)
        return out
    return res.clone()


def all_to_all_vdev_2d(ctx: AllToAllContext, x: torch.Tensor, in_splits: torch.Tensor):
    """Variable splits known only on the device: in_splits int32 [W * g].  Returns (out rows compacted, out_splits [W, g])."""
    cum = torch.zeros(in_splits.numel() + 1, dtype=torch.int32, device=x.device)
    cum[1:] = torch.cumsum(in_splits.to(torch.int32), 0)
    splits, recv, _ = fast_all_to_all(ctx, x.contiguous(), cum)
    return all_to_all_post_process(ctx, splits, recv), splits.clone()


