"""Process-wide expert-parallel op management for the autograd functions (reference: python/triton_dist/function/nvidia/common.py:
``init_triton_dist_ep_op`` / ``init_triton_dist_ep_ctx`` / ``get_triton_dist_ep_op`` / ``deinit_triton_dist_ep_op``, ``MoEOptimConfig``,
the profile switches, ``custom_fwd`` / ``custom_bwd``).

Training frameworks call these once per process and then use the fused EP-MoE function in every layer.  Here the "ep op" is a
:class:`triton_dist.ops.ep_mega.EPMegaContext` (symmetric receive / combine buffers + flags of the Mega-EP kernels); ``split_mbs``
creates two of them so that two micro-batches can be in flight on two streams.
"""
from __future__ import annotations

import functools
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from ...ops import ep_mega as EM

_IMPLS = ("mega", "mega_recomp", "split_mbs")
_STATE: Dict[str, object] = {"ops": [None, None, None], "streams": [None, None, None], "max_tokens": None, "topk": None, "capacity": None}
_PROFILE = {"enabled": False, "output_dir": os.environ.get("TRITON_DIST_PROFILE_DIR", "./triton_dist_profile")}


def custom_fwd(*args, **kwargs):
    """``torch.amp.custom_fwd`` for CUDA (the reference pins ``device_type="cuda"`` the same way)."""
    kwargs.setdefault("device_type", "cuda")
    return torch.amp.custom_fwd(*args, **kwargs)


def custom_bwd(*args, **kwargs):
    kwargs.setdefault("device_type", "cuda")
    return torch.amp.custom_bwd(*args, **kwargs)


def set_triton_dist_moe_profile_enabled(enabled: bool = True, output_dir: Optional[str] = None):
    _PROFILE["enabled"] = bool(enabled)
    if output_dir:
        _PROFILE["output_dir"] = output_dir


def get_triton_dist_moe_profile_enabled() -> dict:
    return dict(_PROFILE)


def get_triton_dist_profile_output_dir() -> str:
    return str(_PROFILE["output_dir"])


def _check_impl(impl: str):
    if impl not in _IMPLS:
        raise ValueError(f"Invalid ep_implementation: {impl}, expected: {list(_IMPLS)}")


class TritonDistEpContext:
    """Per-layer view of the process-wide op: the group, the op (Mega-EP context), its stream and a dict of named events."""

    def __init__(self, ep_group, ep_op: EM.EPMegaContext, ep_stream, ep_events, num_experts_per_rank: int, max_m: int):
        self.ep_group, self.ep_op = ep_group, ep_op
        self.triton_dist_ep_stream, self.triton_dist_ep_events = ep_stream, ep_events if ep_events is not None else {}
        self.num_experts_per_rank, self.max_m = num_experts_per_rank, max_m
        self.max_num_tiles = (max_m + 127) // 128 + num_experts_per_rank       # upper bound of grouped-GEMM row tiles on this rank


def triton_dist_ep_op_initialized(ep_implementation: str = "mega") -> bool:
    _check_impl(ep_implementation)
    ops = _STATE["ops"]
    if ep_implementation == "split_mbs":
        return ops[1] is not None and ops[2] is not None
    return ops[0] is not None


def init_triton_dist_ep_op(ep_group, max_tokens_per_rank: int, hidden_size: int, topk: int, ep_rank: int, num_experts: int, ep_size: int,
                           dtype=torch.bfloat16, weight_dtype=torch.float32, num_sm: int = 8, sm_margin: int = 0, num_buffers: int = 1,
                           capacity: float = 4.0, ep_implementation: str = "mega"):
    """Collective: every rank of the EP group calls it once.  ``num_sm`` = communication CTAs of the Mega-EP kernels; ``weight_dtype``,
    ``sm_margin`` and ``num_buffers`` are accepted for source compatibility (routing weights are fp32, buffers are double-buffered by
    call parity)."""
    _check_impl(ep_implementation)
    import triton_dist.utils as U
    assert ep_size == U.world_size() and ep_rank == U.rank(), "the EP group is the symmetric-heap world"
    cpd = max(1, min(8, int(num_sm) // max(1, ep_size))) if num_sm else 2
    mk = lambda: EM.create_ep_mega_context(max_tokens_per_rank, hidden_size, topk, num_experts, dtype, capacity_factor=float(capacity), cpd=cpd)   # noqa: E731
    cuda = torch.cuda.is_available() and U.current_device().type == "cuda"
    if ep_implementation == "split_mbs":
        _STATE["ops"][1], _STATE["ops"][2] = mk(), mk()
        if cuda:
            _STATE["streams"][1], _STATE["streams"][2] = torch.cuda.Stream(), torch.cuda.Stream()
    else:
        _STATE["ops"][0] = mk()
        if cuda:
            _STATE["streams"][0] = torch.cuda.Stream()
    _STATE.update(max_tokens=max_tokens_per_rank, topk=topk, capacity=float(capacity))
    return get_triton_dist_ep_op(1 if ep_implementation == "split_mbs" else 0)


def deinit_triton_dist_ep_op(ep_implementation: str = "mega"):
    """Collective: frees the symmetric buffers."""
    _check_impl(ep_implementation)
    idxs = (1, 2) if ep_implementation == "split_mbs" else (0,)
    for i in idxs:
        op = _STATE["ops"][i]
        if op is not None:
            op.finalize()
        _STATE["ops"][i] = None
        _STATE["streams"][i] = None


def init_triton_dist_ep_ctx(ep_group, topk: int, num_experts: int, ep_implementation: str = "mega", mbs_idx: int = 0) -> TritonDistEpContext:
    _check_impl(ep_implementation)
    assert triton_dist_ep_op_initialized(ep_implementation), "Please initialize triton_dist_ep_op first."
    size = ep_group.size() if ep_group is not None and hasattr(ep_group, "size") else torch.distributed.get_world_size()
    assert num_experts % size == 0
    idx = (1 + int(mbs_idx)) if ep_implementation == "split_mbs" else 0
    return TritonDistEpContext(ep_group, _STATE["ops"][idx], _STATE["streams"][idx], {}, num_experts // size,
                               int(_STATE["max_tokens"]) * topk * size)


def get_ep_capacity(ep_implementation: str = "mega") -> float:
    assert triton_dist_ep_op_initialized(ep_implementation), "Please initialize triton_dist_ep_op first."
    return float(_STATE["capacity"])


def get_triton_dist_ep_stream(idx: int = 0):
    if idx not in (0, 1, 2):
        raise ValueError(f"Invalid idx: {idx}, expected: [0, 1, 2]")
    return _STATE["streams"][idx]


def get_triton_dist_ep_op(idx: int = 0):
    if idx not in (0, 1, 2):
        raise ValueError(f"Invalid idx: {idx}, expected: [0, 1, 2]")
    return _STATE["ops"][idx]


@dataclass
class MoEOptimConfig:
    """SM / warp budget of the EP kernels.  On B200 the Mega-EP kernels take their communication CTAs from the SMs the grouped GEMM has
    no tiles for; the fields keep the reference's names so launch scripts that tune them keep working."""
    num_build_sms: int
    num_copy_sms: int
    num_group_gemm_warps: int
    num_dispatch_warps: int
    num_combine_warps: int
    num_dispatch_sms: int
    num_tail_sms_in_dispatch: int
    num_combine_sms: int
    num_reduce_sms_in_combine: int
    dispatch_use_block_wise_barrier: bool


@functools.lru_cache(None)
def _sm_count() -> int:
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count if torch.cuda.is_available() else 148


def get_moe_optim_config(use_mega: bool = False, is_forward: bool = True) -> MoEOptimConfig:
    sms = _sm_count()
    comm = 16 if use_mega else 32            # measured sweet spots on 8 x B200: 16 comm CTAs inside the Mega-EP GEMMs, 32 for the split kernels
    return MoEOptimConfig(num_build_sms=8, num_copy_sms=sms, num_group_gemm_warps=8, num_dispatch_warps=8, num_combine_warps=8,
                          num_dispatch_sms=comm, num_tail_sms_in_dispatch=0 if use_mega else 4, num_combine_sms=comm,
                          num_reduce_sms_in_combine=8 if is_forward else 16, dispatch_use_block_wise_barrier=not use_mega)


def fused_ep_moe(x, topk_ids, topk_weights, w_gate_up, w_down, ep_ctx: Optional[TritonDistEpContext] = None):
    """The fused EP-MoE layer on the process-wide op: dispatch || up-projection, SwiGLU, down-projection || combine; differentiable
    (function/nvidia/ep_moe_fused.py)."""
    from .ep_moe_fused import mega_ep_moe_autograd
    op = ep_ctx.ep_op if ep_ctx is not None else get_triton_dist_ep_op(0)
    assert op is not None, "Please initialize triton_dist_ep_op first."
    return mega_ep_moe_autograd(op, x, topk_ids, topk_weights, w_gate_up, w_down)
