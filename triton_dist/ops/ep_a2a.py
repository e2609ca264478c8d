"""Expert-parallel low-latency dispatch / combine (intra-node NVLink), bf16 or online fp8.

Reference API: ``create_ep_ll_a2a_ctx`` + ``dispatch_kernel_v2`` / ``combine_kernel_v2``
(/root/reference/python/triton_dist/kernels/nvidia/low_latency_all_to_all_v2.py:156-696) and the older
``create_all_to_all_context`` / ``fast_all_to_all`` / ``all_to_all_post_process`` (low_latency_all_to_all.py).
Kernels: csrc/ep_kernels.cu.  Output contract:
  dispatch -> recv_x [E/W, W*max_m, H] (fp8 e4m3 or bf16), recv_scale [E/W, W*max_m, H/128] fp32 (fp8 only),
              expert_recv_count [E/W] int32, meta: recv_src_info [E/W, W*max_m] int32 (flat token*topk+k at the source),
              recv_range [E/W, W] int64 = (count << 32 | start)
  combine  -> out [num_tokens, H] = sum_j w[t, j] * expert_out(t, j)
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


class _DispArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("T", c_ll), ("H", c_ll), ("topk", c_ll), ("E", c_ll), ("max_m", c_ll), ("use_fp8", c_ll),
                ("grid", c_ll), ("x", c_void_p), ("topk_idx", c_void_p), ("staging", c_void_p), ("staging_buf_bytes", c_ll),
                ("recv_flag", c_void_p), ("send_count", c_void_p), ("phase", c_void_p), ("recv_x", c_void_p),
                ("recv_scale", c_void_p), ("recv_src_info", c_void_p), ("recv_range", c_void_p), ("recv_count", c_void_p)]


class _CombArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("T", c_ll), ("H", c_ll), ("topk", c_ll), ("E", c_ll), ("max_m", c_ll), ("grid", c_ll),
                ("y", c_void_p), ("src_info", c_void_p), ("recv_range", c_void_p), ("topk_idx", c_void_p), ("topk_w", c_void_p),
                ("comb", c_void_p), ("comb_buf_bytes", c_ll), ("comb_flag", c_void_p), ("phase", c_void_p), ("out", c_void_p)]


_C.register("td_ep_msg_bytes", c_ll, [c_ll, c_int])
_C.register("td_ep_dispatch", c_int, [C.POINTER(_DispArgs), c_void_p])
_C.register("td_ep_combine", c_int, [C.POINTER(_CombArgs), c_void_p])


def _msg_bytes(H: int, fp8: bool, esize: int = 2) -> int:
    payload = H + (H // 128) * 4 if fp8 else H * esize
    return (16 + payload + 15) // 16 * 16


@dataclass
class DispatchMetaInfo:
    recv_token_source_indices: torch.Tensor            # [E/W, W*max_m] int32
    recv_token_source_count_and_start: torch.Tensor    # [E/W, W] int64 (count << 32 | start)

    def counts_and_starts(self) -> Tuple[torch.Tensor, torch.Tensor]:
        v = self.recv_token_source_count_and_start
        return (v >> 32).to(torch.int32), (v & 0xFFFFFFFF).to(torch.int32)


@dataclass
class EPLowLatencyContext:
    max_m: int
    hidden: int
    topk: int
    num_experts: int
    online_quant_fp8: bool
    dtype: torch.dtype
    world_size: int
    rank: int
    staging: torch.Tensor = None        # symmetric uint8 [2, epr, W, max_m, msg]
    recv_flag: torch.Tensor = None      # symmetric int64 [2, epr, W]
    comb: torch.Tensor = None           # symmetric [2, max_m * topk, H]
    comb_flag: torch.Tensor = None      # symmetric int32 [2, W]
    send_count: torch.Tensor = None     # local int32 [E]
    phase_d: torch.Tensor = None        # local int32 [4]
    phase_c: torch.Tensor = None
    host_calls_d: int = 0
    host_calls_c: int = 0

    @property
    def experts_per_rank(self) -> int:
        return self.num_experts // self.world_size

    def finalize(self):
        heap = U.get_heap()
        for t in (self.staging, self.recv_flag, self.comb, self.comb_flag):
            if t is not None:
                heap.free_tensor(t)
        self.staging = self.recv_flag = self.comb = self.comb_flag = None


def create_ep_ll_a2a_ctx(max_m: int, hidden: int, topk: int, num_experts: int, online_quant_fp8: bool = True,
                         fp8_gsize: int = 128, dtype: torch.dtype = torch.bfloat16, world_size: Optional[int] = None,
                         rank: Optional[int] = None) -> EPLowLatencyContext:
    """``max_m``: max tokens a rank dispatches per call (also the per-(expert, source) slot count)."""
    heap = U.get_heap()
    world_size = heap.world if world_size is None else world_size
    rank = heap.rank if rank is None else rank
    assert fp8_gsize == 128 and num_experts % world_size == 0 and hidden % 128 == 0
    ctx = EPLowLatencyContext(max_m, hidden, topk, num_experts, online_quant_fp8, dtype, world_size, rank)
    epr = num_experts // world_size
    esize = torch.empty(0, dtype=dtype).element_size()
    assert esize == 2 or heap.device.type != "cuda", "the CUDA kernels move 16-bit tokens"
    msg = _msg_bytes(hidden, online_quant_fp8, esize)
    ctx.staging = heap.tensor((2, epr, world_size, max_m, msg), torch.uint8)
    ctx.recv_flag = heap.tensor((2, epr, world_size), torch.int64)
    ctx.comb = heap.tensor((2, max_m * topk, hidden), dtype)
    ctx.comb_flag = heap.tensor((2, max(world_size, 4)), torch.int32)
    ctx.send_count = torch.zeros(num_experts, dtype=torch.int32, device=heap.device)
    ctx.phase_d = torch.zeros(4, dtype=torch.int32, device=heap.device)
    ctx.phase_c = torch.zeros(4, dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


def ep_ll_dispatch(ctx: EPLowLatencyContext, x: torch.Tensor, topk_idx: torch.Tensor, num_sms: int = 0):
    """x: [T, H] bf16 (T <= max_m), topk_idx: [T, topk] int32 (-1 = unrouted).
    Returns (recv_x, recv_scale | None, expert_recv_count, DispatchMetaInfo)."""
    T, H = x.shape
    W, epr = ctx.world_size, ctx.experts_per_rank
    assert T <= ctx.max_m and H == ctx.hidden and topk_idx.shape == (T, ctx.topk)
    cap = W * ctx.max_m
    dev = x.device
    fp8 = ctx.online_quant_fp8
    if not x.is_cuda:
        return _dispatch_host(ctx, x, topk_idx)
    recv_x = torch.empty((epr, cap, H), dtype=torch.float8_e4m3fn if fp8 else ctx.dtype, device=dev)
    recv_scale = torch.empty((epr, cap, H // 128), dtype=torch.float32, device=dev) if fp8 else None
    recv_count = torch.empty(epr, dtype=torch.int32, device=dev)
    src_info = torch.empty((epr, cap), dtype=torch.int32, device=dev)
    recv_range = torch.empty((epr, W), dtype=torch.int64, device=dev)
    a = _DispArgs()
    a.symm = symm_args()
    a.T, a.H, a.topk, a.E, a.max_m, a.use_fp8 = T, H, ctx.topk, ctx.num_experts, ctx.max_m, int(fp8)
    a.grid = num_sms or torch.cuda.get_device_properties(dev).multi_processor_count
    a.x, a.topk_idx = x.contiguous().data_ptr(), topk_idx.to(torch.int32).contiguous().data_ptr()
    a.staging, a.staging_buf_bytes = ctx.staging.data_ptr(), ctx.staging[0].numel()
    a.recv_flag, a.send_count, a.phase = ctx.recv_flag.data_ptr(), ctx.send_count.data_ptr(), ctx.phase_d.data_ptr()
    a.recv_x, a.recv_scale = recv_x.data_ptr(), recv_scale.data_ptr() if fp8 else None
    a.recv_src_info, a.recv_range, a.recv_count = src_info.data_ptr(), recv_range.data_ptr(), recv_count.data_ptr()
    _C.check(_C.cuda_lib().td_ep_dispatch(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_ep_dispatch")
    return recv_x, recv_scale, recv_count, DispatchMetaInfo(src_info, recv_range)


def ep_ll_combine(ctx: EPLowLatencyContext, expert_out: torch.Tensor, topk_idx: torch.Tensor, topk_weights: torch.Tensor,
                  meta: DispatchMetaInfo, num_sms: int = 0) -> torch.Tensor:
    """expert_out: [E/W, W*max_m, H] bf16 in the packed dispatch layout -> combined [T, H]."""
    T = topk_idx.shape[0]
    H = ctx.hidden
    if not expert_out.is_cuda:
        return _combine_host(ctx, expert_out, topk_idx, topk_weights, meta)
    out = torch.empty((T, H), dtype=ctx.dtype, device=expert_out.device)
    a = _CombArgs()
    a.symm = symm_args()
    a.T, a.H, a.topk, a.E, a.max_m = T, H, ctx.topk, ctx.num_experts, ctx.max_m
    a.grid = num_sms or torch.cuda.get_device_properties(expert_out.device).multi_processor_count
    a.y, a.src_info, a.recv_range = expert_out.contiguous().data_ptr(), meta.recv_token_source_indices.data_ptr(), meta.recv_token_source_count_and_start.data_ptr()
    a.topk_idx, a.topk_w = topk_idx.to(torch.int32).contiguous().data_ptr(), topk_weights.float().contiguous().data_ptr()
    a.comb, a.comb_buf_bytes, a.comb_flag, a.phase = ctx.comb.data_ptr(), ctx.comb[0].numel() * ctx.comb.element_size(), ctx.comb_flag.data_ptr(), ctx.phase_c.data_ptr()
    a.out = out.data_ptr()
    _C.check(_C.cuda_lib().td_ep_combine(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_ep_combine")
    return out


def last_combined_rows(ctx: EPLowLatencyContext, T: int) -> torch.Tensor:
    """The un-weighted expert outputs ``[T * topk, H]`` that the most recent :func:`ep_ll_combine` received on this rank (row
    ``t * topk + k``).  The weighting happens on the token's owner, so these rows are exactly what the gradient w.r.t. the routing
    weights needs (d w[t, k] = <row(t, k), d out[t]>).  The buffer half comes from the device-resident call counter (no host sync)."""
    n = T * ctx.topk
    if not ctx.comb.is_cuda:
        return ctx.comb[ctx.host_calls_c & 1, :n].clone()
    half = (ctx.phase_c[0:1] & 1).long()
    return torch.index_select(ctx.comb, 0, half)[0, :n].clone()


# reference spellings
def dispatch_kernel_v2(ctx, x, topk_idx, **kw):
    return ep_ll_dispatch(ctx, x, topk_idx, **kw)


def combine_kernel_v2(ctx, expert_out, topk_idx, topk_weights, meta, **kw):
    return ep_ll_combine(ctx, expert_out, topk_idx, topk_weights, meta, **kw)


def dequant_fp8(recv_x: torch.Tensor, recv_scale: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """[..., H] e4m3 + [..., H/128] scales -> dense ``dtype`` (per-128 group scaling)."""
    H = recv_x.shape[-1]
    return (recv_x.float().view(*recv_x.shape[:-1], H // 128, 128) * recv_scale[..., None]).view(*recv_x.shape).to(dtype)


# ------------------------------------------------------------------------------------------------------------
# emulation (no GPU): same protocol (slots, 64-bit phase|count flags, double buffering) with host atomics
# ------------------------------------------------------------------------------------------------------------
def _dispatch_host(ctx, x, topk_idx):
    import ctypes
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, epr = ctx.world_size, ctx.rank, ctx.experts_per_rank
    T, H = x.shape
    ctx.host_calls_d += 1
    ph, par = ctx.host_calls_d, ctx.host_calls_d & 1
    cap = W * ctx.max_m
    msg = ctx.staging.shape[-1]
    counts = [0] * ctx.num_experts
    xb = x.to(ctx.dtype).contiguous().view(torch.uint8).view(T, -1)
    for pair in range(T * ctx.topk):
        e = int(topk_idx.view(-1)[pair])
        if e < 0:
            continue
        dst, le = e // epr, e % epr
        slot = counts[e]
        counts[e] += 1
        m = heap.peer_view(ctx.staging, dst)[par, le, me, slot]
        m[:4] = torch.tensor([pair], dtype=torch.int32).view(torch.uint8)
        m[16:16 + xb.shape[1]] = xb[pair // ctx.topk]
    for e in range(ctx.num_experts):
        dst, le = e // epr, e % epr
        f = heap.peer_ptr(ctx.recv_flag[par, le, me:me + 1].data_ptr(), dst)
        lib.tdh_notify64(ctypes.c_void_p(f), (ph << 32) | counts[e], 1)
    recv_x = torch.zeros((epr, cap, H), dtype=ctx.dtype)
    src_info = torch.zeros((epr, cap), dtype=torch.int32)
    recv_range = torch.zeros((epr, W), dtype=torch.int64)
    recv_count = torch.zeros(epr, dtype=torch.int32)
    for le in range(epr):
        for src in range(W):
            addr = ctx.recv_flag[par, le, src:src + 1].data_ptr()
            while True:
                v = lib.tdh_ld_acquire64(ctypes.c_void_p(addr))
                if (v >> 32) == ph:
                    break
            cnt = v & 0xFFFFFFFF
            start = int(recv_count[le])
            recv_count[le] += cnt
            recv_range[le, src] = (cnt << 32) | start
            for i in range(cnt):
                m = ctx.staging[par, le, src, i]
                src_info[le, start + i] = int(m[:4].view(torch.int32)[0])
                recv_x[le, start + i] = m[16:16 + xb.shape[1]].view(ctx.dtype)
    return recv_x, None, recv_count, DispatchMetaInfo(src_info, recv_range)


def _combine_host(ctx, expert_out, topk_idx, topk_weights, meta):
    import ctypes
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, epr = ctx.world_size, ctx.rank, ctx.experts_per_rank
    ctx.host_calls_c += 1
    ph, par = ctx.host_calls_c, ctx.host_calls_c & 1
    cnts, starts = meta.counts_and_starts()
    for le in range(epr):
        for src in range(W):
            c, s = int(cnts[le, src]), int(starts[le, src])
            dst = heap.peer_view(ctx.comb, src)[par]
            for i in range(c):
                dst[int(meta.recv_token_source_indices[le, s + i])] = expert_out[le, s + i].to(ctx.dtype)
    for r in range(W):
        lib.tdh_notify32(ctypes.c_void_p(heap.peer_ptr(ctx.comb_flag[par, me:me + 1].data_ptr(), r)), ph, 1)
    if lib.tdh_wait32_n(ctypes.c_void_p(ctx.comb_flag[par].data_ptr()), W, ph, 1, 60_000_000):
        raise TimeoutError("ep combine: a peer never returned its rows")
    T = topk_idx.shape[0]
    y = ctx.comb[par, :T * ctx.topk].float().view(T, ctx.topk, -1)
    w = topk_weights.float() * (topk_idx >= 0).float()
    return (y * w[..., None]).sum(1).to(ctx.dtype)
