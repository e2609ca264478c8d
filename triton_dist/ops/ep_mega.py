"""Mega-EP: expert-parallel MoE forward in two fused kernels -- dispatch || grouped GEMM and grouped GEMM || combine.

Reference: kernels/nvidia/ep_all2all_fused.py (tile kernels :73-835, ``mega_dispatch_group_gemm`` :839,
``mega_group_gemm_combine`` :1020) and layers/nvidia/ep_a2a_fused_layer.py:71-763 -- one persistent Triton kernel per half whose
CTAs pull (task type, tile) pairs from a queue, tokens land in a receive buffer and the grouped GEMM gathers them by index.

B200 design (csrc/gemm_sm100.cuh, modes kEPD / kEPC): the per-expert token counts of every rank are all-gathered first (a few
hundred bytes), so EVERY rank can compute where each of its (token, expert) rows belongs in the destination's expert-sorted,
256-row-aligned A matrix.  Half 1 is then ONE kernel: comm CTAs store rows straight into their final position on the
destination (no receive-side index list, sort, or gather -- the tile::gather4 producer measured 3x slower than tiled TMA)
plus a 4-byte return address, and release one flag per (local expert, source, comm CTA); the 2-CTA tcgen05 tiles of an expert
start as soon as its flags are up, experts are sent in the order the GEMM consumes them.  Half 2 is ONE kernel too: the down
projection's epilogue stores every output row directly into its (token, k) slot on the token's owner; the last CTA
release-flags all ranks.  The owner then runs the weighted top-k reduce (one memory-bound kernel over its own rows).
No token saving (a token routed to two experts of one rank is sent twice): rows must sit in their expert's segment for the
tiled TMA path; the extra NVLink bytes are traded for the removed gather / scatter passes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from .. import _C
from .. import utils as U
from . import moe as M
from .elementwise import silu_mul
from .gemm import GemmConfig, fill_common

_TILE = 256


@dataclass
class EPMegaContext:
    max_tokens: int          # tokens per rank per call
    hidden: int
    topk: int
    num_experts: int
    rows_cap: int            # rows this rank can receive (multiple of 256)
    dtype: torch.dtype
    rank: int
    world_size: int
    cpd: int                 # comm CTAs per destination rank
    rx: torch.Tensor = None        # symmetric [2, rows_cap, hidden]
    meta: torch.Tensor = None      # symmetric int32 [2, rows_cap]
    flags: torch.Tensor = None     # symmetric int32 [2, epr, W, cpd]
    comb: torch.Tensor = None      # symmetric [2, max_tokens * topk, hidden]
    done: torch.Tensor = None      # symmetric int32, flat [2][W]
    phase1: torch.Tensor = None    # local int32 [4] (dispatch kernel)
    phase2: torch.Tensor = None    # local int32 [4] (combine kernel)
    calls: int = 0

    @property
    def experts_per_rank(self) -> int:
        return self.num_experts // self.world_size

    def finalize(self):
        heap = U.get_heap()
        if self.rx is None:
            return
        if self.rx.is_cuda:
            torch.cuda.synchronize()
        U.barrier_all_host()
        for t in (self.rx, self.meta, self.flags, self.comb, self.done):
            heap.free_tensor(t)
        self.rx = self.meta = self.flags = self.comb = self.done = None


def create_ep_mega_context(max_tokens: int, hidden: int, topk: int, num_experts: int, dtype: torch.dtype = torch.bfloat16,
                           capacity_factor: float = 2.0, cpd: int = 2) -> EPMegaContext:
    """``capacity_factor``: receive capacity relative to a perfectly balanced routing (every rank receives
    ``max_tokens * topk`` rows); rows beyond it are dropped by the senders."""
    heap = U.get_heap()
    W = heap.world
    assert num_experts % W == 0
    epr = num_experts // W
    rows = int(max_tokens * topk * capacity_factor) + epr * (_TILE - 1)
    rows_cap = (rows + _TILE - 1) // _TILE * _TILE
    assert rows_cap < (1 << 24) and max_tokens * topk < (1 << 24), "return addresses are 24 bits"
    ctx = EPMegaContext(max_tokens, hidden, topk, num_experts, rows_cap, dtype, heap.rank, W, cpd)
    ctx.rx = heap.tensor((2, rows_cap, hidden), dtype)
    ctx.meta = heap.tensor((2, rows_cap), torch.int32)
    ctx.flags = heap.tensor((2, epr, W, cpd), torch.int32)
    ctx.comb = heap.tensor((2, max_tokens * topk, hidden), dtype)
    ctx.done = heap.tensor((max(2 * W, 8),), torch.int32)       # flat [2][W]: done[par * W + s] = call number
    ctx.phase1 = torch.zeros(4, dtype=torch.int32, device=heap.device)
    ctx.phase2 = torch.zeros(4, dtype=torch.int32, device=heap.device)
    U.barrier_all_host()
    return ctx


@dataclass
class EPMegaHandle:
    tile_expert: torch.Tensor     # int32 [rows_cap / 256]: local expert of every 256-row tile, -1 = unused
    route: torch.Tensor           # int32 [rows_cap]: return address of every row (0xffffffff for padding)
    n_rows: torch.Tensor          # int32 [1] rows in use (device)
    parity: int
    T: int
    tot_me: torch.Tensor = None   # int32 [epr]: rows of every local expert
    eoff_me: torch.Tensor = None  # int32 [epr]: first row of every local expert's (256-aligned) segment


def _preprocess(ctx: EPMegaContext, topk_ids: torch.Tensor):
    """Routing bookkeeping from the all-gathered per-expert counts (tiny tensors, device side, no host sync):
    where my rows go on every destination, and how my own A matrix is tiled."""
    W, me, E, epr = ctx.world_size, ctx.rank, ctx.num_experts, ctx.experts_per_rank
    dev = topk_ids.device
    srt = M.moe_align_sort(topk_ids, E, 1)                        # my pairs sorted by global expert (stable)
    send_off = srt.expert_offsets                                 # [E + 1]
    my_cnt = (send_off[1:] - send_off[:-1]).contiguous()
    all_cnt = torch.empty((W, E), dtype=torch.int32, device=dev)
    if W > 1:
        dist.all_gather_into_tensor(all_cnt.view(-1), my_cnt, group=U.get_triton_dist_world())
    else:
        all_cnt[0] = my_cnt
    tot = all_cnt.sum(0, dtype=torch.int32)                       # rows of every expert over all sources
    padded = (tot + (_TILE - 1)) // _TILE * _TILE
    pad2 = padded.view(W, epr)
    eoff = (torch.cumsum(pad2, 1, dtype=torch.int32) - pad2).to(torch.int32)      # segment start of expert e on its rank
    before = (torch.cumsum(all_cnt, 0, dtype=torch.int32) - all_cnt)[me]          # rows of earlier sources, per expert
    dest_off = (eoff.reshape(E) + before).to(torch.int32).contiguous()
    # my own A matrix: tile -> local expert, row -> valid?
    ends = torch.cumsum(pad2[me], 0, dtype=torch.int32)           # [epr]
    n_tiles = ctx.rows_cap // _TILE
    t0 = torch.arange(n_tiles, device=dev, dtype=torch.int32) * _TILE
    te = torch.searchsorted(ends, t0, right=True).to(torch.int32)
    tile_expert = torch.where(t0 < ends[-1], te, torch.full_like(te, -1)).contiguous()
    rows = torch.arange(ctx.rows_cap, device=dev, dtype=torch.int32)
    re = tile_expert.repeat_interleave(_TILE)
    re_c = re.clamp(min=0).long()
    valid = (re >= 0) & ((rows - eoff[me][re_c]) < tot.view(W, epr)[me][re_c])
    return srt, send_off, dest_off, tile_expert, valid, ends[-1:].contiguous(), tot.view(W, epr)[me].contiguous(), eoff[me].contiguous()


def mega_dispatch_group_gemm(ctx: EPMegaContext, x: torch.Tensor, topk_ids: torch.Tensor, w_gate_up: torch.Tensor,
                             config: Optional[GemmConfig] = None):
    """x: ``[T, H]`` my tokens, topk_ids: ``[T, topk]`` global expert ids, w_gate_up: ``[epr, 2I, H]`` my experts.
    -> (``h`` ``[rows_cap, 2I]`` gate/up outputs in my expert-sorted layout, handle).  ONE fused kernel: dispatch || grouped GEMM."""
    W, me, epr = ctx.world_size, ctx.rank, ctx.experts_per_rank
    T, H = x.shape
    assert H == ctx.hidden and T <= ctx.max_tokens and topk_ids.shape == (T, ctx.topk) and x.dtype == ctx.dtype
    E_l, N, K = w_gate_up.shape
    assert E_l == epr and K == H and w_gate_up.is_contiguous()
    srt, send_off, dest_off, tile_expert, valid, n_rows, tot_me, eoff_me = _preprocess(ctx, topk_ids)
    ctx.calls += 1
    par = ctx.calls & 1
    if not x.is_cuda:
        return _dispatch_gemm_host(ctx, x, topk_ids, w_gate_up, srt, send_off, dest_off, tile_expert, valid, n_rows, tot_me, eoff_me, par)
    h = torch.empty((ctx.rows_cap, N), dtype=x.dtype, device=x.device)
    cfg = config or GemmConfig(bn=256 if N % 256 == 0 else 128, cta_group=2, group_m=1, use_tma_store=True)
    xc = x.contiguous()
    args = _C.GemmArgs()
    args.mode = 5
    fill_common(args, ctx.rows_cap, ctx.rx.data_ptr(), H, w_gate_up.reshape(epr * N, K), h.data_ptr(), ctx.rows_cap, h.stride(0),
                ctx.rows_cap, N, K, GemmConfig(cfg.bn, cfg.cta_group, 1, True, 0, W * ctx.cpd), x.dtype == torch.bfloat16)
    args.a_nbuf, args.a_buf_stride_bytes = 2, ctx.rx[0].numel() * ctx.rx.element_size()
    args.tile_expert, args.num_experts = tile_expert.data_ptr(), epr
    r, w_, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = r, w_, base, stride, mc
    args.phase = ctx.phase1.data_ptr()
    args.ag_ws, args.ag_ws_buf_bytes, args.ag_flags = ctx.rx.data_ptr(), args.a_buf_stride_bytes, ctx.flags.data_ptr()
    args.epd_send_off, args.epd_send_ids, args.epd_dest_off = send_off.data_ptr(), srt.sorted_ids.data_ptr(), dest_off.data_ptr()
    args.epd_x, args.epd_topk, args.epd_epr, args.epd_cpd, args.epd_rows_cap = xc.data_ptr(), ctx.topk, epr, ctx.cpd, ctx.rows_cap
    args.epd_meta = ctx.meta.data_ptr()
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemm_launch(mega_ep dispatch)")
    # return addresses of my rows (what the senders wrote), padding rows masked out
    route = torch.where(valid, ctx.meta[par], torch.full_like(ctx.meta[par], -1)).contiguous()
    return h, EPMegaHandle(tile_expert, route, n_rows, par, T, tot_me, eoff_me)


def mega_group_gemm_combine(ctx: EPMegaContext, act: torch.Tensor, handle: EPMegaHandle, w_down: torch.Tensor,
                            topk_weights: torch.Tensor, config: Optional[GemmConfig] = None) -> torch.Tensor:
    """act: ``[rows_cap, I]`` (my sorted layout), w_down: ``[epr, H, I]``, topk_weights ``[T, topk]`` -> ``[T, H]``.
    ONE fused kernel (grouped GEMM whose epilogue delivers every row to its owner) + the owner's weighted top-k reduce."""
    W, epr = ctx.world_size, ctx.experts_per_rank
    E_l, N, K = w_down.shape
    assert E_l == epr and N == ctx.hidden and act.shape == (ctx.rows_cap, K) and w_down.is_contiguous()
    if not act.is_cuda:
        return _gemm_combine_host(ctx, act, handle, w_down, topk_weights)
    cfg = config or GemmConfig(bn=256 if N % 256 == 0 else 128, cta_group=2, group_m=1, use_tma_store=False)
    actc = act.contiguous()
    args = _C.GemmArgs()
    args.mode = 6
    fill_common(args, ctx.rows_cap, actc.data_ptr(), actc.stride(0), w_down.reshape(epr * N, K), ctx.comb.data_ptr(), ctx.rows_cap, N,
                ctx.rows_cap, N, K, GemmConfig(cfg.bn, cfg.cta_group, 1, False, 0, 0), act.dtype == torch.bfloat16)
    args.tile_expert, args.num_experts = handle.tile_expert.data_ptr(), epr
    r, w_, base, stride, mc = U.symm_ctx_fields()
    args.rank, args.world, args.symm_base, args.symm_stride, args.mc_base = r, w_, base, stride, mc
    args.phase = ctx.phase2.data_ptr()
    args.rs_stage, args.rs_stage_buf_bytes = ctx.comb.data_ptr(), ctx.comb[0].numel() * ctx.comb.element_size()
    args.rs_flags, args.c_route = ctx.done.data_ptr(), handle.route.data_ptr()
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemm_launch(mega_ep combine)")
    # owner side: every rank has delivered (done[par][s] == call number), then sum my tokens' top-k rows
    par = handle.parity
    for s in range(W):
        U.wait_eq(ctx.done[par * W + s:par * W + s + 1], ctx.calls, geq=True)
    T = handle.T
    return M.reduce_topk(ctx.comb[par][:T * ctx.topk], topk_weights.to(torch.float32), ctx.topk)


# ------------------------------------------------------------------------------------------------------------
# emulation (no GPU): the SAME protocol on the shared-memory heap -- direct placement from the all-gathered counts, one
# release flag per (local expert, source), return addresses, rows scattered to their owners, per-rank done flags
# ------------------------------------------------------------------------------------------------------------
def _dispatch_gemm_host(ctx, x, topk_ids, w_gate_up, srt, send_off, dest_off, tile_expert, valid, n_rows, tot_me, eoff_me, par):
    import ctypes
    heap, lib = U.get_heap(), _C.host_lib()
    W, me, epr, topk = ctx.world_size, ctx.rank, ctx.experts_per_rank, ctx.topk
    ph = ctx.calls
    timeout = U.get_int_env("TD_HOST_TIMEOUT_US", 60_000_000)
    so, ids_sorted, doff = send_off.tolist(), srt.sorted_ids, dest_off.tolist()
    for e in range(epr):                                     # experts in the order the destination consumes them
        for q in range(W):
            d = (me + q) % W
            g = d * epr + e
            n = so[g + 1] - so[g]
            if n:
                pairs = ids_sorted[so[g]:so[g + 1]].long()
                rows = torch.arange(doff[g], doff[g] + n)
                keep = rows < ctx.rows_cap
                heap.peer_view(ctx.rx, d)[par][rows[keep]] = x[pairs[keep] // topk]
                heap.peer_view(ctx.meta, d)[par][rows[keep]] = ((me << 24) | pairs[keep]).to(torch.int32)
            flag = ctx.flags[par, e, me, 0:1]
            lib.tdh_notify32(ctypes.c_void_p(heap.peer_ptr(flag.data_ptr(), d)), ph, 1)
    N = w_gate_up.shape[1]
    h = torch.zeros((ctx.rows_cap, N), dtype=x.dtype)
    te = tile_expert.tolist()
    seen = set()
    for t, e in enumerate(te):
        if e < 0:
            continue
        if e not in seen:                                    # the tile's expert: all sources have delivered
            for s_ in range(W):
                if lib.tdh_wait32(ctypes.c_void_p(ctx.flags[par, e, s_, 0:1].data_ptr()), ph, 1, timeout):
                    raise TimeoutError(f"mega_ep: rows of expert {e} from rank {s_} never arrived (call {ph})")
            seen.add(e)
        r0 = t * _TILE
        h[r0:r0 + _TILE] = (ctx.rx[par][r0:r0 + _TILE].float() @ w_gate_up[e].float().t()).to(x.dtype)
    route = torch.where(valid, ctx.meta[par], torch.full_like(ctx.meta[par], -1)).contiguous()
    return h, EPMegaHandle(tile_expert, route, n_rows, par, x.shape[0], tot_me, eoff_me)


def _gemm_combine_host(ctx, act, handle, w_down, topk_weights):
    import ctypes
    heap, lib = U.get_heap(), _C.host_lib()
    W, me = ctx.world_size, ctx.rank
    ph, par = ctx.calls, handle.parity
    timeout = U.get_int_env("TD_HOST_TIMEOUT_US", 60_000_000)
    route = handle.route
    for t, e in enumerate(handle.tile_expert.tolist()):
        if e < 0:
            continue
        r0 = t * _TILE
        y = (act[r0:r0 + _TILE].float() @ w_down[e].float().t()).to(act.dtype)
        rt = route[r0:r0 + _TILE]
        for s_ in range(W):                                  # the epilogue's remote scatter: row -> (owner, pair slot)
            m = (rt >= 0) & ((rt >> 24) == s_)
            if m.any():
                heap.peer_view(ctx.comb, s_)[par][(rt[m] & 0xFFFFFF).long()] = y[m]
    for d in range(W):
        lib.tdh_notify32(ctypes.c_void_p(heap.peer_ptr(ctx.done[par * W + me:par * W + me + 1].data_ptr(), d)), ph, 1)
    for s_ in range(W):
        if lib.tdh_wait32(ctypes.c_void_p(ctx.done[par * W + s_:par * W + s_ + 1].data_ptr()), ph, 1, timeout):
            raise TimeoutError(f"mega_ep: rank {s_} never finished its combine stores (call {ph})")
    T = handle.T
    return M.reduce_topk(ctx.comb[par][:T * ctx.topk], topk_weights.to(torch.float32), ctx.topk)


def mega_ep_moe(ctx: EPMegaContext, x, topk_ids, topk_weights, w_gate_up, w_down) -> torch.Tensor:
    """dispatch -> gate/up grouped GEMM -> SwiGLU -> down grouped GEMM -> combine, two fused kernels + the activation."""
    h, handle = mega_dispatch_group_gemm(ctx, x, topk_ids, w_gate_up)
    return mega_group_gemm_combine(ctx, silu_mul(h), handle, w_down, topk_weights)


def mega_ep_moe_reference(x, topk_ids, topk_weights, w_gate_up_all, w_down_all) -> torch.Tensor:
    """fp32 golden with ALL experts' weights (``[E, 2I, H]``, ``[E, H, I]``): what the distributed op must reproduce."""
    T, H = x.shape
    out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
    xf = x.float()
    for k in range(topk_ids.shape[1]):
        for e in topk_ids[:, k].unique().tolist():
            if e < 0:
                continue
            m = topk_ids[:, k] == e
            hh = xf[m] @ w_gate_up_all[e].float().t()
            I = hh.shape[1] // 2
            a = (torch.nn.functional.silu(hh[:, :I]) * hh[:, I:]).to(x.dtype).float()
            out[m] += topk_weights[m, k:k + 1].float() * (a @ w_down_all[e].float().t()).to(x.dtype).float()
    return out
