"""Tensor-parallel MoE block: experts' FFN dimension sharded over TP ranks.

Reference: /root/reference/python/triton_dist/layers/nvidia/tp_moe.py:237-276 -- router (cuBLAS) -> softmax/top-k ->
NCCL all_gather of ids & weights -> ag_group_gemm -> silu*mul -> run_moe_reduce_rs.  Here the router GEMM, the token
all-gather, the grouped GEMMs and the reduce-scatter all run on our kernels; the tiny id/weight all-gather uses the
low-latency push all-gather instead of NCCL.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import utils as U
from ..ops import comm
from ..ops import moe as M
from ..ops.elementwise import silu_mul
from .tp_mlp import _linear


class TP_MoE:
    def __init__(self, rank: int = 0, world_size: int = 8, group=None):
        self.rank, self.world_size, self.group = rank, world_size, group
        self.router = None          # [E, H] replicated
        self.w_gate_up = None       # [E, 2 * I / W, H]   (K-major)
        self.w_down = None          # [E, H, I / W]       (K-major)
        self.ag_ctx = self.rs_ctx = self.meta_ag = None
        self.topk = 2
        self.norm_topk_prob = True

    def _init_parameters_from_shards(self, router, w_gate_up, w_down, topk: int, norm_topk_prob: bool = True):
        self.router, self.w_gate_up, self.w_down = router, w_gate_up, w_down
        self.num_experts, self.topk, self.norm_topk_prob = router.shape[0], topk, norm_topk_prob
        self.hidden = router.shape[1]
        self.dtype = w_gate_up.dtype

    def _init_parameters(self, moe, verbose: bool = False):
        """HF Qwen3MoeSparseMoeBlock-like module: ``gate`` (router) + ``experts[i].{gate,up,down}_proj``."""
        dev = U.current_device()
        W, r = self.world_size, self.rank
        gu, dn = [], []
        for ex in moe.experts:
            g = ex.gate_proj.weight.detach().chunk(W, 0)[r]
            u = ex.up_proj.weight.detach().chunk(W, 0)[r]
            gu.append(torch.cat((g, u), 0))
            dn.append(ex.down_proj.weight.detach().chunk(W, 1)[r])
        self._init_parameters_from_shards(moe.gate.weight.detach().to(dev), torch.stack(gu).to(dev).contiguous(),
                                          torch.stack(dn).to(dev).contiguous(), moe.top_k, getattr(moe, "norm_topk_prob", True))

    def _init_ctx(self, max_M: int):
        E, I2, H = self.w_gate_up.shape
        self.ag_ctx = M.create_ag_group_gemm_context(max_M, I2, H, E, self.topk, self.dtype, self.rank, self.world_size)
        self.rs_ctx = M.create_moe_rs_context(self.rank, self.world_size, self.world_size, max_M * self.topk, H, E, self.topk, self.dtype)
        self.meta_ag = comm.create_fast_allgather_context(max(1024, (max_M // self.world_size) * self.topk * 8), self.rank, self.world_size)

    def finalize(self):
        for c in (self.ag_ctx, self.rs_ctx, self.meta_ag):
            if c is not None:
                c.finalize()
        self.ag_ctx = self.rs_ctx = self.meta_ag = None

    def _route(self, x2: torch.Tensor):
        logits = _linear(x2, self.router).float()
        probs = torch.softmax(logits, dim=-1)
        w, ids = torch.topk(probs, self.topk, dim=-1)
        if self.norm_topk_prob:
            w = w / w.sum(-1, keepdim=True)
        return ids.to(torch.int32), w

    @torch.inference_mode()
    def torch_fwd(self, x: torch.Tensor) -> torch.Tensor:
        """Baseline on replicated activations: per-expert matmuls (cuBLAS) + NCCL all-reduce."""
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        ids, w = self._route(x2)
        T = x2.shape[0]
        out = torch.zeros((T, self.hidden), dtype=torch.float32, device=x.device)
        for e in range(self.num_experts):
            tok, k = torch.where(ids == e)
            if tok.numel() == 0:
                continue
            h = torch.nn.functional.linear(x2[tok], self.w_gate_up[e])
            I = h.shape[1] // 2
            h = torch.nn.functional.silu(h[:, :I]) * h[:, I:]
            y = torch.nn.functional.linear(h, self.w_down[e]).float()
            out.index_add_(0, tok, y * w[tok, k][:, None])
        out = out.to(x.dtype)
        if self.world_size > 1:
            dist.all_reduce(out, group=self.group)
        return out.view(shp)

    @torch.inference_mode()
    def dist_triton_fwd(self, x: torch.Tensor) -> torch.Tensor:
        """``x``: this rank's tokens ``[T/W, H]`` -> ``[T/W, H]``  (AG-MoE up-proj, MoE-reduce-RS down-proj)."""
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        W = self.world_size
        ids, w = self._route(x2)
        if W > 1:
            full_ids = comm.fast_allgather(ids.contiguous(), self.meta_ag, mode="push").view(-1, self.topk)
            full_w = comm.fast_allgather(w.contiguous(), self.meta_ag, mode="push").view(-1, self.topk)
        else:
            full_ids, full_w = ids, w
        h = M.ag_group_gemm(x2, self.w_gate_up, self.ag_ctx, full_ids)          # [T * topk, 2I/W]
        h = silu_mul(h)
        out = M.run_moe_reduce_rs(h, self.w_down, full_ids, full_w, self.rs_ctx)   # [T/W, H]
        return out.view(shp)

    @torch.inference_mode()
    def dist_triton_AR_fwd(self, x: torch.Tensor) -> torch.Tensor:
        """Replicated activations: local grouped GEMMs + fast all-reduce (moe_reduce_ar)."""
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        ids, w = self._route(x2)
        h = M.moe_forward_local(x2, self.w_gate_up, ids)
        h = silu_mul(h)
        out = M.run_moe_reduce_ar(h, self.w_down, ids, w, self.rs_ctx)
        return out.view(shp)

    dist_triton_gemm_ar_fwd = dist_triton_AR_fwd


def shard_local(t: torch.Tensor, world_size: int, dim: int, local_rank: int) -> torch.Tensor:
    """This rank's slice of ``t`` along ``dim`` (reference: layers/nvidia/tp_moe.py ``shard_local``)."""
    assert t.shape[dim] % world_size == 0
    n = t.shape[dim] // world_size
    return t.narrow(dim, local_rank * n, n).contiguous()
