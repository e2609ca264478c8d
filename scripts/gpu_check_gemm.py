"""First-contact check of the tcgen05 GEMM on a real B200: numerics per config + TFLOPS vs cuBLAS.
Usage: python scripts/gpu_check_gemm.py <cta_group> [perf]
Each invocation is wrapped in `timeout` by the caller so a hung config cannot take the whole call down."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from triton_dist.ops import GemmConfig, gemm  # noqa: E402

cg = int(sys.argv[1])
perf = len(sys.argv) > 2
res = []


def run(M, N, K, bn, tma, dtype=torch.bfloat16, group_m=4):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda", dtype=dtype)
    b = torch.randn(N, K, device="cuda", dtype=dtype)
    c = gemm(a, b, config=GemmConfig(bn=bn, cta_group=cg, group_m=group_m, use_tma_store=tma))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (c.float() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    r = dict(M=M, N=N, K=K, bn=bn, cg=cg, tma=tma, dtype=str(dtype), max_abs_err=err, rel=rel, ok=rel < 2e-2)
    print(json.dumps(r), flush=True)
    res.append(r)


def bench(M, N, K, bn, tma=True, group_m=8, iters=20):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cfg = GemmConfig(bn=bn, cta_group=cg, group_m=group_m, use_tma_store=tma)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        gemm(a, b, out=c, config=cfg)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gemm(a, b, out=c, config=cfg); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ours = sorted(ts)[len(ts) // 2]
    ts = []
    bt = b.t()
    for _ in range(3):
        torch.matmul(a, bt, out=c)
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, bt, out=c); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    cub = sorted(ts)[len(ts) // 2]
    fl = 2.0 * M * N * K
    r = dict(perf=True, M=M, N=N, K=K, bn=bn, cg=cg, tma=tma, group_m=group_m, ours_ms=ours, cublas_ms=cub,
             ours_tflops=fl / ours / 1e9, cublas_tflops=fl / cub / 1e9)
    print(json.dumps(r), flush=True)
    res.append(r)


if not perf:
    run(128 * cg, 256, 64, 256, False)          # one tile, one k-block
    run(128 * cg, 256, 256, 256, False)         # several k-blocks (pipeline wrap)
    run(128 * cg, 256, 1024, 256, False)        # > stages k-blocks
    run(512, 768, 512, 256, False)
    run(512, 768, 512, 256, True)
    for bn in (32, 64, 128):
        run(512, 768, 512, bn, False)
        run(512, 768, 512, bn, True)
    run(77, 1000, 520, 128, False)
    run(300, 264, 1032, 64, True)
    run(4096, 4096, 4096, 256, True, group_m=8)
    run(2048, 2048, 2048, 128, True, torch.float16)
else:
    for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (4096, 12288, 6144), (8192, 1536, 4096), (8192, 4096, 1536)]:
        for bn in (256, 128):
            for gm in (8, 16):
                bench(M, N, K, bn, True, gm)
    bench(8192, 8192, 8192, 256, False, 8)
    for bn in (32, 64, 128):
        bench(16, 4096, 4096, bn, False, 1)
        bench(128, 12288, 4096, bn, False, 1)
import os
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/gemm_check_cg{cg}{'_perf' if perf else ''}.json", "w") as f:
    json.dump(res, f, indent=1)
