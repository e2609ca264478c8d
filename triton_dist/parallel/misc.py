"""Small layer wrappers kept for API parity: ``GemmARLayer`` (layers/nvidia/gemm_allreduce_layer.py:143) and
``AllGatherLayer`` (layers/nvidia/low_latency_allgather_layer.py:197)."""
from __future__ import annotations

import torch

from .. import utils as U
from ..ops import comm
from ..ops.gemm_ar import create_gemm_ar_context, create_ll_gemm_ar_context, gemm_allreduce_op, low_latency_gemm_allreduce_op


class GemmARLayer:
    def __init__(self, tp_group, max_M: int, N: int, K: int, input_dtype: torch.dtype, output_dtype: torch.dtype,
                 local_world_size: int, persistent: bool = True, use_ll_kernel: bool = False, copy_to_local: bool = True,
                 NUM_COMM_SMS: int = 16):
        heap = U.get_heap()
        self.use_ll_kernel = use_ll_kernel
        make = create_ll_gemm_ar_context if use_ll_kernel else create_gemm_ar_context
        self.ctx = make(None, heap.rank, heap.world, local_world_size, max_M, N, output_dtype, NUM_COMM_SMS=NUM_COMM_SMS)

    def forward(self, x: torch.Tensor, weight: torch.Tensor, bias=None, scale_a=None, scale_b=None) -> torch.Tensor:
        """16-bit ``x`` / ``weight``: fused GEMM + AllReduce.  int8 / float8_e4m3fn operands: ``scale_a`` (per-tensor or per-row) and
        ``scale_b`` (per-tensor or per-output-channel) are applied in the GEMM epilogue (reference gemm_allreduce_layer.py:143)."""
        if x.dtype in (torch.bfloat16, torch.float16, torch.float32):
            if scale_a is not None or scale_b is not None:
                raise ValueError("GemmARLayer: scale_a / scale_b are dequantisation scales of int8 / float8_e4m3fn operands")
            out = (low_latency_gemm_allreduce_op if self.use_ll_kernel else gemm_allreduce_op)(self.ctx, x, weight)
        else:
            out = gemm_allreduce_op(self.ctx, x, weight, As=scale_a, Bs=scale_b)
        return out if bias is None else out + bias

    __call__ = forward

    def finalize(self):
        self.ctx.finalize()


class AllGatherLayer:
    """Low-latency all-gather layer (reference: layers/nvidia/low_latency_allgather_layer.py:33-140).  Five kernels behind the
    reference's method names: pull, push, push-LL (flag-in-data), and the two NVLS forms (multimem push, multimem LL).  The staged
    variants of the reference (push_3d, push_numa_2d*) relay through one GPU per NUMA node / node because PCIe or inter-node links
    are their bottleneck; inside one NVSwitch domain every peer is one hop away, so those names run the direct kernel of the same
    protocol family (LL stays LL).  ``stages`` is accepted for source compatibility: buffers are double-buffered by call parity."""

    def __init__(self, max_shard_bytes: int, stages: int = 2):
        self.ctx = comm.create_fast_allgather_context(max_shard_bytes)
        self.stages = stages

    def _fwd(self, x, mode):
        return comm.fast_allgather(x, self.ctx, mode=mode)

    def forward_pull(self, x): return self._fwd(x, "pull")
    def forward_push_2d(self, x): return self._fwd(x, "push")
    def forward_push_3d(self, x): return self._fwd(x, "push")
    def forward_push_2d_ll(self, x): return self._fwd(x, "push_2d_ll")
    def forward_push_numa_2d(self, x): return self._fwd(x, "push")
    def forward_push_numa_2d_ll(self, x): return self._fwd(x, "push_2d_ll")
    def forward_push_multimem(self, x): return self._fwd(x, "push_multimem")
    def forward_push_2d_ll_multimem(self, x): return self._fwd(x, "push_2d_ll_multimem")

    def forward(self, x, mode: str = "auto"):
        """auto: LL for tiny shards (no barrier round-trip), NVLS push for the rest when the multicast mapping exists."""
        if mode == "auto":
            nbytes = x.numel() * x.element_size()
            mc = U.is_nvshmem_multimem_supported()
            mode = ("push_2d_ll_multimem" if mc else "push_2d_ll") if nbytes <= (8 << 10) else ("push_multimem" if mc else "push")
        return self._fwd(x, mode)

    __call__ = forward

    def finalize(self):
        self.ctx.finalize()
