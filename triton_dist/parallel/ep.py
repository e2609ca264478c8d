"""Expert-parallel layers on the low-latency dispatch/combine kernels.

Reference: layers/nvidia/ep_ll_a2a_layer.py (EPLowLatencyAllToAllLayer), ep_a2a_layer.py (EPConfig,
DispatchCombineContext, EPAll2AllLayer), ep_moe.py (EP_MoE), ep_a2a_fused_layer.py (EpAll2AllFusedOp) and
function/nvidia/ep_moe_fused.py (TritonDistFusedEpMoeFunction: fwd + bwd).

Forward = route -> dispatch (NVLink push, optional online fp8) -> grouped tcgen05 GEMMs over the packed per-expert
rows -> combine (NVLink push + weighted top-k sum).  Backward reuses the same two collectives with roles swapped
(grad of combine is a dispatch of the output gradient, grad of dispatch is a combine of the row gradients) and
computes wgrad per expert.
"""
from __future__ import annotations

import dataclasses

from dataclasses import dataclass
from typing import Optional

import torch

from .. import utils as U
from ..ops import ep_a2a as EP
from ..ops import moe as M
from ..ops.elementwise import silu_mul
from ..ops.gemm import GemmConfig
from .tp_mlp import _linear


@dataclass
class EPConfig:
    max_tokens: int
    hidden: int
    topk: int
    num_experts: int
    rank: int
    world_size: int
    online_quant_fp8: bool = False
    dtype: torch.dtype = torch.bfloat16


class EPLowLatencyAllToAllLayer:
    def __init__(self, max_m: int, hidden: int, topk: int, num_experts: int, online_quant_fp8: bool = True,
                 rank: Optional[int] = None, world_size: Optional[int] = None, dtype: torch.dtype = torch.bfloat16):
        self.ctx = EP.create_ep_ll_a2a_ctx(max_m, hidden, topk, num_experts, online_quant_fp8, 128, dtype, world_size, rank)

    def dispatch(self, send_tokens: torch.Tensor, send_scale, topk_indices: torch.Tensor):
        return EP.ep_ll_dispatch(self.ctx, send_tokens, topk_indices)

    def combine(self, expert_out: torch.Tensor, topk_indices: torch.Tensor, topk_weights: torch.Tensor, meta):
        return EP.ep_ll_combine(self.ctx, expert_out, topk_indices, topk_weights, meta)

    # ---- tracing (reference: dump_dispatch_trace / dump_combine_trace, ep_ll_a2a_layer.py) ------------------
    def _dump_trace(self, name: str, fn, path: str, iters: int = 5):
        """Device-timed spans of ``iters`` calls written as a Chrome / Perfetto trace (one track per rank)."""
        import json
        spans = []
        if torch.cuda.is_available() and self.ctx.staging.is_cuda:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            origin = torch.cuda.Event(enable_timing=True)
            origin.record()
            for a, b in evs:
                a.record(); fn(); b.record()
            torch.cuda.synchronize()
            spans = [(origin.elapsed_time(a) * 1e3, a.elapsed_time(b) * 1e3) for a, b in evs]
        else:
            import time
            t0 = time.perf_counter()
            for _ in range(iters):
                s0 = time.perf_counter(); fn(); spans.append(((s0 - t0) * 1e6, (time.perf_counter() - s0) * 1e6))
        ev = [{"name": name, "ph": "X", "ts": ts, "dur": dur, "pid": f"rank{self.ctx.rank}", "tid": name} for ts, dur in spans]
        with open(path, "w") as f:
            json.dump({"traceEvents": ev}, f)
        return spans

    def dump_dispatch_trace(self, send_tokens: torch.Tensor, topk_indices: torch.Tensor, path: str = "ep_dispatch_trace.json", iters: int = 5):
        return self._dump_trace("ep_ll_dispatch", lambda: self.dispatch(send_tokens, None, topk_indices), path, iters)

    def dump_combine_trace(self, expert_out, topk_indices, topk_weights, meta, path: str = "ep_combine_trace.json", iters: int = 5):
        return self._dump_trace("ep_ll_combine", lambda: self.combine(expert_out, topk_indices, topk_weights, meta), path, iters)

    def finalize(self):
        self.ctx.finalize()


class EPNormalAll2AllLayer:
    """Throughput-mode EP layer (token saving; reference layers/nvidia/ep_a2a_layer.py: preprocess / dispatch /
    dispatch_postprocess / combine) on ops/ep_normal.py.  ``dispatch`` returns the handle the expert FFN consumes through
    index lists (no receive-side copy); ``combine`` pre-reduces per token on the expert rank."""

    def __init__(self, ep_config: "EPConfig"):
        from ..ops import ep_normal as EN
        self.cfg, self.EN = ep_config, EN
        self.ctx = EN.create_ep_normal_ctx(ep_config.max_tokens, ep_config.hidden, ep_config.topk, ep_config.num_experts, ep_config.dtype)

    def preprocess(self, topk_indices: torch.Tensor):
        return M.histogram_by_expert(topk_indices, self.cfg.num_experts)

    def dispatch(self, x: torch.Tensor, topk_indices: torch.Tensor, topk_weights: torch.Tensor):
        return self.EN.ep_dispatch_normal(self.ctx, x, topk_indices, topk_weights)

    def expert_ffn(self, handle, w_gate_up: torch.Tensor, w_down: torch.Tensor) -> torch.Tensor:
        return self.EN.ep_expert_ffn_normal(self.ctx, handle, w_gate_up, w_down)

    def combine(self, y_pairs: torch.Tensor, handle, topk_indices: torch.Tensor) -> torch.Tensor:
        return self.EN.ep_combine_normal(self.ctx, y_pairs, handle, topk_indices)

    def finalize(self):
        self.ctx.finalize()


class EPAll2AllLayer(EPLowLatencyAllToAllLayer):
    """Throughput ("normal") mode: same protocol, bf16 payload, capacity sized for prefill-scale token counts."""

    def __init__(self, ep_config: EPConfig):
        super().__init__(ep_config.max_tokens, ep_config.hidden, ep_config.topk, ep_config.num_experts, False,
                         ep_config.rank, ep_config.world_size, ep_config.dtype)
        self.cfg = ep_config

    def preprocess(self, topk_indices: torch.Tensor):
        return M.histogram_by_expert(topk_indices, self.cfg.num_experts)

    def dispatch_postprocess(self, recv, meta):
        return recv


def packed_tile_experts(counts: torch.Tensor, cap: int, block_m: int = 128) -> torch.Tensor:
    """tile -> local expert id for the packed dispatch layout [epr, cap, H] (cap % block_m == 0); -1 = empty tile."""
    epr = counts.numel()
    tiles_per = cap // block_m
    t = torch.arange(tiles_per, device=counts.device)[None, :] * block_m
    e = torch.arange(epr, device=counts.device, dtype=torch.int32)[:, None].expand(epr, tiles_per)
    return torch.where(t < counts[:, None], e, torch.full_like(e, -1)).reshape(-1).contiguous()


def grouped_ffn_packed(x_packed: torch.Tensor, counts: torch.Tensor, w_gate_up: torch.Tensor, w_down: torch.Tensor) -> torch.Tensor:
    """SwiGLU FFN over the packed rows: x_packed [epr, cap, H] -> [epr, cap, H]; rows >= count are don't-care."""
    epr, cap, H = x_packed.shape
    te = packed_tile_experts(counts, cap, 128) if cap % 128 == 0 else None
    if te is None:
        out = torch.zeros_like(x_packed)
        for le in range(epr):
            n = int(counts[le])
            h = _linear(x_packed[le, :n].contiguous(), w_gate_up[le]) if n else x_packed[le, :0]
            out[le, :n] = _linear(silu_mul(h), w_down[le]) if n else out[le, :0]
        return out
    r = M.SortedRouting(None, te, None, None, epr * cap, 128, -1)
    h = M.moe_grouped_gemm(x_packed.view(epr * cap, H), w_gate_up, r)
    h = silu_mul(h)
    y = M.moe_grouped_gemm(h, w_down, r)
    return y.view(epr, cap, H)


class EP_MoE:
    """Experts sharded over ranks (``E / W`` local experts with full FFN width)."""

    def __init__(self, rank: int = 0, world_size: int = 8, group=None):
        self.rank, self.world_size, self.group = rank, world_size, group
        self.router = self.w_gate_up = self.w_down = None
        self.a2a: Optional[EPLowLatencyAllToAllLayer] = None

    def _init_parameters_from_shards(self, router, w_gate_up, w_down, topk: int, norm_topk_prob: bool = True):
        """router [E, H] (replicated); w_gate_up [E/W, 2I, H]; w_down [E/W, H, I]."""
        self.router, self.w_gate_up, self.w_down = router, w_gate_up, w_down
        self.num_experts, self.topk, self.norm_topk_prob = router.shape[0], topk, norm_topk_prob
        self.hidden, self.dtype = router.shape[1], w_gate_up.dtype

    def _init_ctx(self, max_tokens: int, online_quant_fp8: bool = False, mode: str = "low_latency"):
        """``mode``: "low_latency" (per-(token, k) messages, optional fp8, packed per-expert receive layout) or "normal"
        (throughput mode: token saving, index-list receive side, local pre-reduce on combine)."""
        max_m = (max_tokens + 127) // 128 * 128          # packed layout stays tile aligned
        self.mode = mode
        if mode == "normal":
            self.a2a = EPNormalAll2AllLayer(EPConfig(max_m, self.hidden, self.topk, self.num_experts, self.rank, self.world_size,
                                                     False, self.dtype))
            return
        self.a2a = EPLowLatencyAllToAllLayer(max_m, self.hidden, self.topk, self.num_experts, online_quant_fp8, self.rank,
                                             self.world_size, self.dtype)

    def finalize(self):
        if self.a2a is not None:
            self.a2a.finalize()
            self.a2a = None

    def _route(self, x2):
        probs = torch.softmax(_linear(x2, self.router).float(), dim=-1)
        w, ids = torch.topk(probs, self.topk, dim=-1)
        if self.norm_topk_prob:
            w = w / w.sum(-1, keepdim=True)
        return ids.to(torch.int32), w

    @torch.inference_mode()
    def torch_fwd(self, x: torch.Tensor) -> torch.Tensor:
        """Golden: every rank evaluates its local experts on the all-gathered tokens, NCCL all_to_all-free formulation."""
        import torch.distributed as dist
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        ids, w = self._route(x2)
        W, epr = self.world_size, self.num_experts // self.world_size
        T = x2.shape[0]
        if W > 1:
            xs = torch.empty((W * T, shp[-1]), dtype=x.dtype, device=x.device)
            dist.all_gather_into_tensor(xs.view(-1), x2.contiguous().view(-1), group=self.group)
            idl = torch.empty((W * T, self.topk), dtype=torch.int32, device=x.device)
            dist.all_gather_into_tensor(idl.view(-1), ids.contiguous().view(-1), group=self.group)
            wl = torch.empty((W * T, self.topk), dtype=torch.float32, device=x.device)
            dist.all_gather_into_tensor(wl.view(-1), w.float().contiguous().view(-1), group=self.group)
        else:
            xs, idl, wl = x2, ids, w.float()
        out = torch.zeros((W * T, shp[-1]), dtype=torch.float32, device=x.device)
        for le in range(epr):
            e = self.rank * epr + le
            tok, k = torch.where(idl == e)
            if tok.numel() == 0:
                continue
            h = torch.nn.functional.linear(xs[tok], self.w_gate_up[le])
            I = h.shape[1] // 2
            y = torch.nn.functional.linear(torch.nn.functional.silu(h[:, :I]) * h[:, I:], self.w_down[le]).float()
            out.index_add_(0, tok, y * wl[tok, k][:, None])
        if W > 1:
            dist.all_reduce(out, group=self.group)
        return out[self.rank * T:(self.rank + 1) * T].to(x.dtype).view(shp)

    @torch.inference_mode()
    def dist_triton_fwd(self, x: torch.Tensor) -> torch.Tensor:
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        ids, w = self._route(x2)
        if getattr(self, "mode", "low_latency") == "normal":
            h = self.a2a.dispatch(x2, ids, w.float())
            y = self.a2a.expert_ffn(h, self.w_gate_up, self.w_down)
            return self.a2a.combine(y, h, ids).view(shp)
        act, handle = self.dispatch_group_gemm(x2, ids, w)
        return self.group_gemm_combine(act, handle).view(shp)

    # the two halves the reference exposes as Mega-EP entry points (ep_all2all_fused.py:839, :1020)
    def dispatch_group_gemm(self, x2: torch.Tensor, ids: torch.Tensor, w: torch.Tensor):
        """dispatch -> grouped GEMM (gate|up) -> SwiGLU on the packed ``[epr, cap, *]`` layout."""
        rx, rs, cnt, meta = self.a2a.dispatch(x2, None, ids)
        xin = EP.dequant_fp8(rx, rs, self.dtype) if rs is not None else rx
        epr, cap, H = xin.shape
        if cap % 128:
            return xin, (ids, w, meta, cnt, None)
        r = M.SortedRouting(None, packed_tile_experts(cnt, cap, 128), None, None, epr * cap, 128, -1)
        act = silu_mul(M.moe_grouped_gemm(xin.view(epr * cap, H), self.w_gate_up, r))
        return act, (ids, w, meta, cnt, r)

    def group_gemm_combine(self, act: torch.Tensor, handle) -> torch.Tensor:
        """grouped GEMM (down) -> combine (weighted top-k reduce at the token's source rank)."""
        ids, w, meta, cnt, r = handle
        if r is None:
            y = grouped_ffn_packed(act, cnt, self.w_gate_up, self.w_down)
        else:
            epr = cnt.numel()
            y = M.moe_grouped_gemm(act, self.w_down, r).view(epr, -1, self.hidden)
        return self.a2a.combine(y, ids, w, meta)


# ------------------------------------------------------------------------------------------------------------
# training path: autograd function over dispatch -> experts -> combine  (function/nvidia/ep_moe_fused.py:42-359)
# ------------------------------------------------------------------------------------------------------------
class _Dispatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ids, layer: EPLowLatencyAllToAllLayer):
        rx, rs, cnt, meta = layer.dispatch(x.detach().contiguous(), None, ids)
        ctx.layer, ctx.ids, ctx.meta, ctx.T = layer, ids, meta, x.shape[0]
        ctx.mark_non_differentiable(cnt)
        return rx, cnt, meta.recv_token_source_indices, meta.recv_token_source_count_and_start

    @staticmethod
    def backward(ctx, g_rx, _gc, _gi, _gr):
        # gradient of "copy row t to each of its experts" = sum over k of the row gradients = unweighted combine
        ones = torch.ones(ctx.ids.shape, dtype=torch.float32, device=g_rx.device)
        gx = ctx.layer.combine(g_rx.contiguous().to(ctx.layer.ctx.dtype), ctx.ids, ones, ctx.meta)
        return gx, None, None


class _Combine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, ids, w, src_info, recv_range, layer: EPLowLatencyAllToAllLayer):
        meta = EP.DispatchMetaInfo(src_info, recv_range)
        out = layer.combine(y.detach().contiguous(), ids, w.detach(), meta)
        ctx.layer, ctx.meta = layer, meta
        # un-weighted rows as they arrived on this rank: needed for the gradient of the routing weights
        rows = EP.last_combined_rows(layer.ctx, ids.shape[0]) if w.requires_grad else None
        ctx.save_for_backward(y, ids, w, rows)
        return out

    @staticmethod
    def backward(ctx, g_out):
        y, ids, w, rows = ctx.saved_tensors
        layer = ctx.layer
        # d y_row(t,k) = w[t,k] * g_out[t]: deliver g_out rows with the forward routing, scale on the expert side with
        # the (all-gathered) routing weights
        g_rows, _, _, meta2 = layer.dispatch(g_out.contiguous().to(layer.ctx.dtype), None, ids)
        W = layer.ctx.world_size
        import torch.distributed as dist
        w_all = torch.empty((W,) + tuple(w.shape), dtype=torch.float32, device=w.device)
        if W > 1:
            dist.all_gather_into_tensor(w_all.view(-1), w.float().contiguous().view(-1), group=U.get_triton_dist_world())
        else:
            w_all[0] = w.float()
        cnts, starts = meta2.counts_and_starts()
        scale = torch.zeros(g_rows.shape[:2], dtype=torch.float32, device=w.device)
        for le in range(g_rows.shape[0]):
            for src in range(W):
                c, s = int(cnts[le, src]), int(starts[le, src])
                if c:
                    scale[le, s:s + c] = w_all[src].reshape(-1)[meta2.recv_token_source_indices[le, s:s + c].long()]
        g_y = (g_rows.float() * scale[..., None]).to(y.dtype)
        # d w[t, k] = <y_row(t, k), g_out[t]>: the un-weighted rows were delivered to this rank by the forward combine
        g_w = None
        if rows is not None:
            T, topk = ids.shape
            g_w = (rows.float().view(T, topk, -1) * g_out.float()[:, None, :]).sum(-1)
            g_w = torch.where(ids >= 0, g_w, torch.zeros_like(g_w)).to(w.dtype)
        return g_y, None, g_w, None, None, None


class TritonDistFusedEpMoeFunction(torch.autograd.Function):
    """Inference-grade fused forward; gradients w.r.t. the token activations flow through dispatch/combine.  Expert
    weights receive gradients through ordinary autograd of the packed per-expert matmuls."""

    @staticmethod
    def apply_moe(x, ids, w, layer: EPLowLatencyAllToAllLayer, w_gate_up: torch.Tensor, w_down: torch.Tensor):
        rx, cnt, src_info, recv_range = _Dispatch.apply(x, ids, layer)
        epr, cap, H = rx.shape
        valid = torch.arange(cap, device=rx.device)[None, :] < cnt[:, None]
        h = torch.einsum("ech,eih->eci", rx.float(), w_gate_up.float())
        I = h.shape[-1] // 2
        h = torch.nn.functional.silu(h[..., :I]) * h[..., I:]
        y = torch.einsum("eci,ehi->ech", h, w_down.float()) * valid[..., None]
        return _Combine.apply(y.to(rx.dtype), ids, w, src_info, recv_range, layer)


class EpAll2AllFusedOp:
    """The reference's Mega-EP op (layers/nvidia/ep_a2a_fused_layer.py:71-763): lazily sized symmetric buffers, then
    ``mega_dispatch_group_gemm`` (dispatch fused with the gate/up grouped GEMM) and ``mega_group_gemm_combine`` (down grouped GEMM
    fused with the combine transfer).  Each half is ONE kernel here (ops/ep_mega.py, csrc/gemm_sm100.cuh modes kEPD / kEPC): comm
    CTAs store every routed row straight into its final expert-sorted position on the destination while the tcgen05 tiles of
    already-complete experts run; the down projection's epilogue delivers every output row to its (token, k) slot on the owner."""

    def __init__(self, ep_config: EPConfig, capacity_factor: float = 2.0):
        self.cfg = ep_config
        self.capacity_factor = capacity_factor
        self.ctx = None

    # lazy allocation (the reference sizes the NVSHMEM heap from these numbers before materialising)
    def get_nvshmem_size(self) -> int:
        c = self.cfg
        esz = torch.empty(0, dtype=c.dtype).element_size()
        rows = int(c.max_tokens * c.topk * self.capacity_factor) + (c.num_experts // c.world_size) * 255
        return 2 * (rows * c.hidden * esz + rows * 4) + 2 * c.max_tokens * c.topk * c.hidden * esz + 4096

    get_nvshmem_size_gb = lambda self: self.get_nvshmem_size() / 2 ** 30

    def materialize(self):
        if U.get_heap().device.type != "cuda":
            self._host_layer()
            return self
        if self.ctx is None:
            from ..ops import ep_mega as EM
            c = self.cfg
            self.ctx = EM.create_ep_mega_context(c.max_tokens, c.hidden, c.topk, c.num_experts, c.dtype, self.capacity_factor)
        return self

    def preprocess(self, topk_indices: torch.Tensor):
        return M.histogram_by_expert(topk_indices, self.cfg.num_experts)

    def mega_dispatch_group_gemm(self, x: torch.Tensor, topk_indices: torch.Tensor, topk_weights: torch.Tensor, w_gate_up: torch.Tensor):
        """-> (SwiGLU activations in my expert-sorted layout ``[rows_cap, I]``, handle)."""
        from ..ops import ep_mega as EM
        from ..ops.elementwise import silu_mul
        self.materialize()
        if not x.is_cuda:
            return self._host_dispatch(x, topk_indices, topk_weights, w_gate_up)
        h, handle = EM.mega_dispatch_group_gemm(self.ctx, x, topk_indices, w_gate_up)
        return silu_mul(h), (handle, topk_weights)

    mega_preprocess_group_gemm = mega_dispatch_group_gemm

    def mega_group_gemm_combine(self, act: torch.Tensor, handle, w_down: torch.Tensor) -> torch.Tensor:
        from ..ops import ep_mega as EM
        if not act.is_cuda:
            return self._host_combine(act, handle, w_down)
        hd, topk_weights = handle
        return EM.mega_group_gemm_combine(self.ctx, act, hd, w_down, topk_weights)

    # emulation backend: the throughput-mode exchange executes the same dispatch / combine protocol on the shared-memory heap
    def _host_layer(self):
        if getattr(self, "_layer", None) is None:
            self._layer = EPNormalAll2AllLayer(self.cfg)
        return self._layer

    def _host_dispatch(self, x, topk_indices, topk_weights, w_gate_up):
        from ..ops import ep_normal as EN
        layer = self._host_layer()
        h = layer.dispatch(x, topk_indices, topk_weights)
        act, r = EN.ep_ffn_up_normal(layer.ctx, h, w_gate_up)
        return act, (h, r, topk_indices)

    def _host_combine(self, act, handle, w_down):
        from ..ops import ep_normal as EN
        h, r, topk_indices = handle
        layer = self._host_layer()
        y = EN.ep_ffn_down_normal(layer.ctx, h, act, r, w_down)
        return layer.combine(y, h, topk_indices)

    def finalize(self):
        if self.ctx is not None:
            self.ctx.finalize()
            self.ctx = None
        if getattr(self, "_layer", None) is not None:
            self._layer.finalize()
            self._layer = None


@dataclasses.dataclass
class EPAllToAllLayoutDesc:
    """What a dispatch produced, for the combine that follows (reference: layers/nvidia/ep_a2a_layer.py ``EPAllToAllLayoutDesc``):
    per-expert receive counts of this rank, the routing of the local tokens, and the kernel's own handle."""
    num_dispatch_token_cur_rank: Optional[torch.Tensor] = None      # int32 [experts_per_rank]: rows received per local expert
    topk_indices: Optional[torch.Tensor] = None                     # [T, topk] routing of MY tokens
    topk_weights: Optional[torch.Tensor] = None
    handle: object = None                                           # DispatchMetaInfo (low latency) / EPNormalHandle (throughput mode)


DispatchCombineContext = EPAllToAllLayoutDesc


def prepare_moe_metadata_using_kernel(topk_ids: torch.Tensor, num_experts: int, block_m: int = 128):
    """Routing metadata for a grouped GEMM in one sort kernel (reference: layers/nvidia/ep_moe.py): (sorted pair ids, tile -> expert map,
    padded per-expert row offsets [E + 1])."""
    from ..ops import moe as _M
    r = _M.moe_align_sort(topk_ids, num_experts, block_m)
    return r.sorted_ids, r.tile_expert, r.expert_offsets
