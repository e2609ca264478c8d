"""Tutorial 07 -- MoE expert GEMMs without gather / scatter passes: the grouped tcgen05 GEMM fetches each tile's token rows with TMA
``tile::gather4`` and writes every output row to its final place from the epilogue; with TD_MOE_AG_FUSED=1 the all-gather of the
tokens is fused into the same kernel (reference: allgather_group_gemm.py, moe_utils.py).  Needs a B200.
    python tutorials/07_moe_tma_gather.py"""
import torch
from triton_dist.ops import moe as M

T, E, topk, K, N = 2048, 32, 4, 2048, 1024
x = (torch.randn(T, K, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(E, N, K, device="cuda") * 0.05).to(torch.bfloat16)
ids = torch.stack([torch.randperm(E, device="cuda")[:topk] for _ in range(T)]).to(torch.int32)
r = M.moe_align_sort(ids, E, 128)                       # expert-sorted, tile-aligned routing (one CUDA kernel)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


staged = lambda: M.scatter_rows(M.moe_grouped_gemm(M.gather_rows(x, r, div=topk), w, r), r, T * topk)
fused = lambda: M.moe_grouped_gemm_fused(x, w, r, topk, T * topk)
assert torch.equal(staged(), fused())
flops = 2.0 * T * topk * N * K
for name, fn in (("gather -> GEMM -> scatter (3 kernels)", staged), ("TMA gather4 GEMM + scatter epilogue (1 kernel)", fused)):
    ms = timed(fn)
    print(f"{name}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.0f} TFLOP/s")
