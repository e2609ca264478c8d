"""Expert-parallel dispatch / combine, throughput ("normal") mode with token saving; kernels: csrc/ep_normal_kernels.cu.

Reference: ``kernel_dispatch_token_intra_node`` / ``kernel_combine_token_intra_node`` (kernels/nvidia/ep_a2a_intra_node.py:39-289),
``ep_dispatch_token_inplace`` / ``ep_combine_token_inplace`` (ep_a2a.py:881,962), ``EPAll2AllLayer`` (layers/nvidia/ep_a2a_layer.py) and
the dispatch + grouped-GEMM / grouped-GEMM + combine halves of the Mega-EP op (ep_all2all_fused.py:839,1020).

What is different from the low-latency path (ops/ep_a2a.py):
  * a token whose experts live on the same rank travels ONCE to that rank (+ one 16-byte descriptor per expert);
  * received rows stay where they landed (``rx[src][slot]``): :func:`ep_expert_ffn_normal` feeds the expert GEMM through an index
    list with TMA ``tile::gather4`` -- no compaction into a per-expert layout;
  * combine pre-reduces a token's expert outputs on the expert rank and returns one row per (token, rank).

STATUS: validated on 2xB200 (tests/dist_worker.py::case_ep_normal: token-saving counts, FFN through the TMA-gather4 index lists,
pre-reduced combine) and on the emulation backend; not yet timed.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _C
from .. import utils as U
from . import moe as M
from .comm import SymmArgs, symm_args

c_void_p, c_ll, c_int = C.c_void_p, C.c_longlong, C.c_int


class _DispArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("T", c_ll), ("H", c_ll), ("topk", c_ll), ("E", c_ll), ("T_max", c_ll), ("P_max", c_ll), ("grid", c_ll),
                ("x", c_void_p), ("topk_idx", c_void_p), ("topk_w", c_void_p),
                ("rx", c_void_p), ("rx_buf_bytes", c_ll), ("rp", c_void_p), ("rmeta", c_void_p), ("rflag", c_void_p),
                ("send_rows", c_void_p), ("send_pairs", c_void_p), ("phase", c_void_p),
                ("pair_expert", c_void_p), ("pair_row", c_void_p), ("rcnt", c_void_p)]


class _CombArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("T", c_ll), ("H", c_ll), ("topk", c_ll), ("E", c_ll), ("T_max", c_ll), ("P_max", c_ll), ("grid", c_ll),
                ("y", c_void_p), ("rp", c_void_p), ("rmeta", c_void_p), ("rcnt", c_void_p), ("topk_idx", c_void_p),
                ("comb", c_void_p), ("comb_buf_bytes", c_ll), ("comb_flag", c_void_p), ("phase", c_void_p), ("out", c_void_p)]


_C.register("td_ep_dispatch_normal", c_int, [C.POINTER(_DispArgs), c_void_p])
_C.register("td_ep_combine_normal", c_int, [C.POINTER(_CombArgs), c_void_p])


@dataclass
class EPNormalContext:
    T_max: int
    hidden: int
    topk: int
    num_experts: int
    dtype: torch.dtype
    rank: int
    world_size: int
    rx: torch.Tensor = None          # symmetric [2, W, T_max, H]
    rp: torch.Tensor = None          # symmetric int32 [2, W, P_max, 4]
    rmeta: torch.Tensor = None       # symmetric int32 [2, W, T_max, 4]
    rflag: torch.Tensor = None       # symmetric int64 [2, W, 2]
    comb: torch.Tensor = None        # symmetric [2, W, T_max, H]
    comb_flag: torch.Tensor = None   # symmetric int32 [2, W]
    send_rows: torch.Tensor = None
    send_pairs: torch.Tensor = None
    phase_d: torch.Tensor = None
    phase_c: torch.Tensor = None
    host_calls_d: int = 0
    host_calls_c: int = 0
    grid: int = 64

    @property
    def P_max(self) -> int:
        return self.T_max * self.topk

    @property
    def experts_per_rank(self) -> int:
        return self.num_experts // self.world_size

    def finalize(self):
        heap = U.get_heap()
        for t in (self.rx, self.rp, self.rmeta, self.rflag, self.comb, self.comb_flag):
            if t is not None:
                heap.free_tensor(t)
        self.rx = self.rp = self.rmeta = self.rflag = self.comb = self.comb_flag = None


def create_ep_normal_ctx(max_tokens: int, hidden: int, topk: int, num_experts: int, dtype: torch.dtype = torch.bfloat16,
                         rank: Optional[int] = None, world_size: Optional[int] = None) -> EPNormalContext:
    heap = U.get_heap()
    rank = heap.rank if rank is None else rank
    W = heap.world if world_size is None else world_size
    ctx = EPNormalContext(max_tokens, hidden, topk, num_experts, dtype, rank, W)
    ctx.rx = heap.tensor((2, W, max_tokens, hidden), dtype)
    ctx.rp = heap.tensor((2, W, ctx.P_max, 4), torch.int32)
    ctx.rmeta = heap.tensor((2, W, max_tokens, 4), torch.int32)
    ctx.rflag = heap.tensor((2, W, 2), torch.int64)
    ctx.comb = heap.tensor((2, W, max_tokens, hidden), dtype)
    ctx.comb_flag = heap.tensor((2, W), torch.int32)
    dev = heap.device
    ctx.send_rows = torch.zeros(max(W, 4), dtype=torch.int32, device=dev)
    ctx.send_pairs = torch.zeros(max(W, 4), dtype=torch.int32, device=dev)
    ctx.phase_d = torch.zeros(4, dtype=torch.int32, device=dev)
    ctx.phase_c = torch.zeros(4, dtype=torch.int32, device=dev)
    U.barrier_all_host()
    return ctx


@dataclass
class EPNormalHandle:
    par: int                    # buffer half of this dispatch
    rx_flat: torch.Tensor       # [W * T_max, H] received payload rows (in place)
    pair_expert: torch.Tensor   # int32 [W * P_max] local expert of every received (token, k) pair, -1 = empty
    pair_row: torch.Tensor      # int32 [W * P_max] row of rx_flat, -1 = empty
    rcnt: torch.Tensor          # int32 [W, 2] rows / pairs per source
    T: int


def ep_dispatch_normal(ctx: EPNormalContext, x: torch.Tensor, topk_idx: torch.Tensor, topk_w: torch.Tensor) -> EPNormalHandle:
    """x: [T, H]; topk_idx: int32 [T, topk] (global expert ids, < 0 = unrouted); topk_w: fp32 [T, topk]."""
    T, H = x.shape
    assert H == ctx.hidden and T <= ctx.T_max and topk_idx.shape == (T, ctx.topk)
    W = ctx.world_size
    if not x.is_cuda:
        return _dispatch_host(ctx, x, topk_idx, topk_w)
    dev = x.device
    pair_expert = torch.empty(W * ctx.P_max, dtype=torch.int32, device=dev)
    pair_row = torch.empty(W * ctx.P_max, dtype=torch.int32, device=dev)
    rcnt = torch.empty((W, 2), dtype=torch.int32, device=dev)
    par = (ctx.host_calls_d + 1) & 1
    a = _DispArgs()
    a.symm = symm_args()
    a.T, a.H, a.topk, a.E, a.T_max, a.P_max, a.grid = T, H, ctx.topk, ctx.num_experts, ctx.T_max, ctx.P_max, ctx.grid
    xc, ic, wc = x.contiguous(), topk_idx.to(torch.int32).contiguous(), topk_w.float().contiguous()
    a.x, a.topk_idx, a.topk_w = xc.data_ptr(), ic.data_ptr(), wc.data_ptr()
    a.rx, a.rx_buf_bytes = ctx.rx.data_ptr(), W * ctx.T_max * H * x.element_size()
    a.rp, a.rmeta, a.rflag = ctx.rp.data_ptr(), ctx.rmeta.data_ptr(), ctx.rflag.data_ptr()
    a.send_rows, a.send_pairs, a.phase = ctx.send_rows.data_ptr(), ctx.send_pairs.data_ptr(), ctx.phase_d.data_ptr()
    a.pair_expert, a.pair_row, a.rcnt = pair_expert.data_ptr(), pair_row.data_ptr(), rcnt.data_ptr()
    _C.check(_C.cuda_lib().td_ep_dispatch_normal(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_ep_dispatch_normal")
    ctx.host_calls_d += 1
    return EPNormalHandle(par, ctx.rx[par].view(W * ctx.T_max, H), pair_expert, pair_row, rcnt, T)


def ep_ffn_up_normal(ctx: EPNormalContext, h: EPNormalHandle, w_gate_up: torch.Tensor):
    """Gate/up grouped GEMM + SwiGLU over the received pairs.  Returns ``(act [capacity, I] in the expert-sorted layout, routing)``.
    On GPUs the GEMM gathers its rows straight out of ``rx`` with TMA tile::gather4 (index list = received row of every sorted pair)."""
    from .elementwise import silu_mul
    r = M.moe_align_sort(h.pair_expert.view(-1, 1), ctx.experts_per_rank, 128)
    n_pairs = h.pair_expert.numel()
    if h.rx_flat.is_cuda:
        valid = r.sorted_ids != r.pad_id
        g = torch.where(valid, h.pair_row[r.sorted_ids.clamp(max=n_pairs - 1).long()], torch.full_like(r.sorted_ids, -1))
        hid = M.moe_grouped_gemm_fused(h.rx_flat, w_gate_up, r, 1, r.capacity, gather_idx=g, scatter=False)
    else:
        ids = r.sorted_ids.long()
        valid = ids != r.pad_id
        xs = torch.zeros((r.capacity, h.rx_flat.shape[1]), dtype=h.rx_flat.dtype)
        xs[valid] = h.rx_flat[h.pair_row[ids[valid]].long()]
        hid = M.moe_grouped_gemm(xs, w_gate_up, r)
    return silu_mul(hid), r


def ep_ffn_down_normal(ctx: EPNormalContext, h: EPNormalHandle, act: torch.Tensor, r, w_down: torch.Tensor) -> torch.Tensor:
    """Down grouped GEMM; rows leave the epilogue in received-pair order ``[W * P_max, H]`` (what combine consumes)."""
    n_pairs = h.pair_expert.numel()
    if act.is_cuda:
        return M.moe_grouped_gemm_fused(act, w_down, r, 1, n_pairs, gather=False)
    ys = M.moe_grouped_gemm(act, w_down, r)
    ids = r.sorted_ids.long()
    valid = ids != r.pad_id
    y = torch.zeros((n_pairs, ys.shape[1]), dtype=ys.dtype)
    y[ids[valid]] = ys[valid]
    return y


def ep_expert_ffn_normal(ctx: EPNormalContext, h: EPNormalHandle, w_gate_up: torch.Tensor, w_down: torch.Tensor) -> torch.Tensor:
    """SwiGLU FFN of the local experts over the received pairs; returns ``y_pairs`` [W * P_max, H] in received-pair order.
    w_gate_up: [epr, 2I, H], w_down: [epr, H, I].  The two halves are the dispatch + grouped GEMM / grouped GEMM + combine stages of the
    reference's Mega-EP op (:func:`ep_ffn_up_normal`, :func:`ep_ffn_down_normal`)."""
    act, r = ep_ffn_up_normal(ctx, h, w_gate_up)
    return ep_ffn_down_normal(ctx, h, act, r, w_down)


def ep_combine_normal(ctx: EPNormalContext, y_pairs: torch.Tensor, h: EPNormalHandle, topk_idx: torch.Tensor) -> torch.Tensor:
    """y_pairs: [W * P_max, H] expert outputs in received-pair order -> [T, H] on the tokens' owner."""
    W, H = ctx.world_size, ctx.hidden
    if not y_pairs.is_cuda:
        return _combine_host(ctx, y_pairs, h, topk_idx)
    out = torch.empty((h.T, H), dtype=y_pairs.dtype, device=y_pairs.device)
    a = _CombArgs()
    a.symm = symm_args()
    a.T, a.H, a.topk, a.E, a.T_max, a.P_max, a.grid = h.T, H, ctx.topk, ctx.num_experts, ctx.T_max, ctx.P_max, ctx.grid
    yc, ic = y_pairs.contiguous(), topk_idx.to(torch.int32).contiguous()
    a.y, a.rp, a.rmeta, a.rcnt, a.topk_idx = yc.data_ptr(), ctx.rp[h.par].data_ptr(), ctx.rmeta[h.par].data_ptr(), h.rcnt.data_ptr(), ic.data_ptr()
    a.comb, a.comb_buf_bytes = ctx.comb.data_ptr(), W * ctx.T_max * H * y_pairs.element_size()
    a.comb_flag, a.phase, a.out = ctx.comb_flag.data_ptr(), ctx.phase_c.data_ptr(), out.data_ptr()
    _C.check(_C.cuda_lib().td_ep_combine_normal(C.byref(a), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_ep_combine_normal")
    ctx.host_calls_c += 1
    return out


# ------------------------------------------------------------------------------------------------------------
# emulation (no GPU): the same buffers, descriptors and flags on the shared-memory heap
# ------------------------------------------------------------------------------------------------------------
def _dispatch_host(ctx, x, topk_idx, topk_w):
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, epr = ctx.world_size, ctx.rank, ctx.experts_per_rank
    T, H = x.shape
    ctx.host_calls_d += 1
    ph, par = ctx.host_calls_d, ctx.host_calls_d & 1
    rows, pairs = [0] * W, [0] * W
    xd = x.to(ctx.dtype)
    for t in range(T):
        ids = [int(v) for v in topk_idx[t]]
        by_dst = {}
        for k, e in enumerate(ids):
            if 0 <= e < ctx.num_experts:
                by_dst.setdefault(e // epr, []).append(k)
        for d, ks in by_dst.items():
            slot, pbase = rows[d], pairs[d]
            rows[d] += 1
            pairs[d] += len(ks)
            heap.peer_view(ctx.rx, d)[par, me, slot] = xd[t]
            heap.peer_view(ctx.rmeta, d)[par, me, slot] = torch.tensor([t, pbase, len(ks), 0], dtype=torch.int32)
            for r_, k in enumerate(ks):
                wbits = int(torch.tensor([float(topk_w[t, k])], dtype=torch.float32).view(torch.int32)[0])
                heap.peer_view(ctx.rp, d)[par, me, pbase + r_] = torch.tensor([slot, ids[k] % epr, t * ctx.topk + k, wbits], dtype=torch.int32)
    for d in range(W):
        f = ctx.rflag[par, me]
        lib.tdh_notify64(C.c_void_p(heap.peer_ptr(f[0:1].data_ptr(), d)), (ph << 32) | rows[d], 1)
        lib.tdh_notify64(C.c_void_p(heap.peer_ptr(f[1:2].data_ptr(), d)), (ph << 32) | pairs[d], 1)
    pair_expert = torch.full((W * ctx.P_max,), -1, dtype=torch.int32)
    pair_row = torch.full((W * ctx.P_max,), -1, dtype=torch.int32)
    rcnt = torch.zeros((W, 2), dtype=torch.int32)
    for src in range(W):
        vals = []
        for j in range(2):
            addr = ctx.rflag[par, src, j:j + 1].data_ptr()
            while True:
                v = lib.tdh_ld_acquire64(C.c_void_p(addr))
                if (v >> 32) == ph:
                    break
            vals.append(v & 0xFFFFFFFF)
        rcnt[src, 0], rcnt[src, 1] = vals
        n = vals[1]
        ent = ctx.rp[par, src, :n]
        pair_expert[src * ctx.P_max:src * ctx.P_max + n] = ent[:, 1]
        pair_row[src * ctx.P_max:src * ctx.P_max + n] = src * ctx.T_max + ent[:, 0]
    return EPNormalHandle(par, ctx.rx[par].view(W * ctx.T_max, H), pair_expert, pair_row, rcnt, T)


def _combine_host(ctx, y_pairs, h, topk_idx):
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, epr, H = ctx.world_size, ctx.rank, ctx.experts_per_rank, ctx.hidden
    ctx.host_calls_c += 1
    ph, par = ctx.host_calls_c, ctx.host_calls_c & 1
    rp, rmeta = ctx.rp[h.par], ctx.rmeta[h.par]
    for src in range(W):
        for slot in range(int(h.rcnt[src, 0])):
            t, first, n = (int(v) for v in rmeta[src, slot, :3])
            acc = torch.zeros(H, dtype=torch.float32)
            for i in range(n):
                w = float(rp[src, first + i, 3:4].view(torch.float32)[0])
                acc += w * y_pairs[src * ctx.P_max + first + i].float()
            heap.peer_view(ctx.comb, src)[par, me, t] = acc.to(ctx.dtype)
    for d in range(W):
        lib.tdh_notify32(C.c_void_p(heap.peer_ptr(ctx.comb_flag[par, me:me + 1].data_ptr(), d)), ph, 1)
    if lib.tdh_wait32_n(C.c_void_p(ctx.comb_flag[par].data_ptr()), W, ph, 1, U.get_int_env("TD_HOST_TIMEOUT_US", 60_000_000)):
        raise TimeoutError("ep_combine_normal: a peer never returned its rows")
    out = torch.zeros((h.T, H), dtype=torch.float32)
    for t in range(h.T):
        for d in sorted({int(e) // epr for e in topk_idx[t] if 0 <= int(e) < ctx.num_experts}):
            out[t] += ctx.comb[par, d, t].float()
    return out.to(ctx.dtype)
