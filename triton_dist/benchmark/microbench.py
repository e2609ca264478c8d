"""Device micro-benchmarks (the role of the reference's little_kernel/benchmark/{memory,compute,latency} tree): the measured
denominators our rooflines use.  Single GPU:  python -m triton_dist.benchmark.microbench
  * HBM copy bandwidth            (our 16-byte vector copy kernel vs cudaMemcpy D2D)
  * L2-resident copy bandwidth    (working set 32 MiB << 126 MB L2)
  * tcgen05 GEMM throughput       (bf16 2-CTA 256x256 tiles, MXFP8 256-wide tiles) vs cuBLAS
  * kernel launch + tiny-kernel latency (signal kernel), CUDA-graph replay latency
"""
import json

import torch

from triton_dist.ops import comm
from triton_dist.ops.gemm import GemmConfig, gemm


def _time(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda")
    res = {}
    for name, nbytes in (("hbm", 1 << 30), ("l2", 32 << 20)):
        src = torch.empty(nbytes, dtype=torch.uint8, device=dev).random_(0, 255)
        dst = torch.empty_like(src)
        ms = _time(lambda: comm.copy_tensor(dst, src))
        res[f"{name}_copy_ours_GBps"] = round(2 * nbytes / ms / 1e6, 1)
        ms = _time(lambda: dst.copy_(src))
        res[f"{name}_copy_memcpy_GBps"] = round(2 * nbytes / ms / 1e6, 1)
    for (M, N, K) in ((8192, 8192, 8192), (4096, 12288, 6144)):
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = _time(lambda: gemm(a, b, out=out, config=GemmConfig(256, 2, 8, True)))
        res[f"gemm_bf16_{M}x{N}x{K}_TFLOPs"] = round(2 * M * N * K / ms / 1e9, 1)
        ms = _time(lambda: torch.matmul(a, b.t(), out=out))
        res[f"cublas_bf16_{M}x{N}x{K}_TFLOPs"] = round(2 * M * N * K / ms / 1e9, 1)
        try:
            from triton_dist.ops.fp8 import gemm_mxfp8, quantize_mxfp8
            qa, qb = quantize_mxfp8(a), quantize_mxfp8(b)
            ms = _time(lambda: gemm_mxfp8(qa, qb, out=out, config=GemmConfig(256, 2, 8, True)))
            res[f"gemm_mxfp8_{M}x{N}x{K}_TFLOPs"] = round(2 * M * N * K / ms / 1e9, 1)
        except Exception as e:      # noqa: BLE001
            res["mxfp8_error"] = str(e)[:100]
    flag = torch.zeros(8, dtype=torch.int32, device=dev)
    from triton_dist import language as dl
    ms = _time(lambda: flag.add_(1), iters=200)
    res["tiny_kernel_launch_us"] = round(ms * 1e3, 2)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            flag.add_(1)
    ms = _time(g.replay, iters=50)
    res["graph_node_us"] = round(ms * 1e3 / 20, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
