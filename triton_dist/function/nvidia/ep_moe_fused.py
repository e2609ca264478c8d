"""Expert-parallel MoE with a full training path (forward + backward) on the Mega-EP kernels.

Reference: python/triton_dist/function/nvidia/ep_moe_fused.py:42-359 (``TritonDistFusedEpMoeFunction``; backward at :186: dgrad
through dispatch + grouped GEMM, wgrad through ``transposed_moe_grouped_gemm``, kernels/nvidia/group_gemm.py:503-727,988).

Every heavy step is one of this repo's kernels (csrc/gemm_sm100.cuh):
  forward : [dispatch || gate/up grouped GEMM] (kEPD) -> SwiGLU kernel -> [down grouped GEMM || combine] (kEPC) -> top-k reduce
  backward: [dispatch of dOut || grouped GEMM with W_down^T] (kEPD) gives d(act) rows directly in the expert-sorted layout ->
            SwiGLU backward kernel -> [grouped GEMM with W_gate_up^T || combine] (kEPC) returns dX to the token owners;
            the two weight gradients are ONE segmented-K batch launch each (``transposed_moe_grouped_gemm``).
No Python loop over experts, no host sync: the routing bookkeeping stays on the device.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ... import utils as U
from ...ops import ep_mega as EM
from ...ops import moe as M
from ...ops.elementwise import silu_mul, silu_mul_backward


class MegaEpMoeFunction(torch.autograd.Function):
    """``out[t] = sum_k w[t, k] * FFN_{e(t, k)}(x[t])`` over experts sharded across ranks; differentiable in x, the routing
    weights and both expert weight tensors."""

    @staticmethod
    def forward(ctx, x, topk_ids, topk_w, w_gate_up, w_down, mctx: EM.EPMegaContext):
        h, handle = EM.mega_dispatch_group_gemm(mctx, x, topk_ids, w_gate_up)
        act = silu_mul(h)
        par = handle.parity
        need_w = topk_w.requires_grad
        out = EM.mega_group_gemm_combine(mctx, act, handle, w_down, topk_w)
        T = x.shape[0]
        # rows as they arrived (input of the gate/up GEMM) and, for d(routing weights), the un-weighted expert outputs per pair
        x_sorted = mctx.rx[par].clone()
        y_pairs = mctx.comb[par][:T * mctx.topk].clone() if need_w else None
        ctx.mctx, ctx.handle = mctx, handle
        ctx.save_for_backward(x, topk_ids, topk_w, w_gate_up, w_down, h, act, x_sorted, y_pairs if y_pairs is not None else torch.empty(0))
        ctx.need_w = need_w
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, ids, w, w_gu, w_dn, h, act, x_sorted, y_pairs = ctx.saved_tensors
        mctx, handle = ctx.mctx, ctx.handle
        W, epr, topk = mctx.world_size, mctx.experts_per_rank, mctx.topk
        T = x.shape[0]
        g_out = g_out.contiguous().to(x.dtype)
        # (1) dOut rows travel with the forward routing; the same kernel multiplies them with W_down^T:  d(act)[r] = dY[r] @ W_down[e]
        w_dn_t = w_dn.transpose(1, 2).contiguous()                      # [epr, I, H]
        d_act_u, handle2 = EM.mega_dispatch_group_gemm(mctx, g_out, ids, w_dn_t)
        # routing weight of every received row: weights of all ranks (tiny all-gather), indexed by the row's return address
        w_all = torch.empty((W, T * topk), dtype=torch.float32, device=x.device)
        if W > 1:
            dist.all_gather_into_tensor(w_all.view(-1), w.float().contiguous().view(-1), group=U.get_triton_dist_world())
        else:
            w_all[0] = w.float().view(-1)
        route = handle2.route
        valid = route >= 0
        src, pair = (route >> 24).clamp(min=0).long(), (route & 0xFFFFFF).long()
        row_w = torch.where(valid, w_all[src, pair.clamp(max=T * topk - 1)], torch.zeros((), device=x.device))
        d_act = (d_act_u.float() * row_w[:, None]).to(x.dtype)
        dy_sorted = (mctx.rx[handle2.parity].float() * row_w[:, None]).to(x.dtype)      # weighted dOut rows in my sorted layout
        # (2) SwiGLU backward, (3) weight gradients: one segmented-K batch launch each
        d_h = silu_mul_backward(d_act, h)
        tot, eoff = handle2.tot_me, handle2.eoff_me
        g_w_dn = M.transposed_moe_grouped_gemm(dy_sorted, act, tot, eoff + tot)        # [epr, H, I]
        g_w_gu = M.transposed_moe_grouped_gemm(d_h, x_sorted, tot, eoff + tot)         # [epr, 2I, H]
        # (4) dX: grouped GEMM with W_gate_up^T, rows returned to their owners and summed over k (un-weighted)
        w_gu_t = w_gu.transpose(1, 2).contiguous()                      # [epr, H, 2I]
        ones = torch.ones((T, topk), dtype=torch.float32, device=x.device)
        g_x = EM.mega_group_gemm_combine(mctx, d_h, handle2, w_gu_t, ones)
        g_w = None
        if ctx.need_w:
            g_w = (y_pairs.float().view(T, topk, -1) * g_out.float()[:, None, :]).sum(-1)
            g_w = torch.where(ids >= 0, g_w, torch.zeros_like(g_w)).to(w.dtype)
        return g_x, None, g_w, g_w_gu, g_w_dn, None


def mega_ep_moe_autograd(mctx: EM.EPMegaContext, x, topk_ids, topk_w, w_gate_up, w_down):
    return MegaEpMoeFunction.apply(x, topk_ids, topk_w, w_gate_up, w_down, mctx)


def __getattr__(name):
    # the reference's class name for the EP-MoE autograd function; the layer-level implementation lives with the EP layers
    if name == "TritonDistFusedEpMoeFunction":
        from ...parallel.ep import TritonDistFusedEpMoeFunction
        return TritonDistFusedEpMoeFunction
    raise AttributeError(name)

