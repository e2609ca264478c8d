#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fp8_gpu.py -x -q 2>&1 | tail -15
timeout 200 python - <<'PY' 2>&1 | tail -12
import sys, torch, json
sys.path.insert(0, ".")
from triton_dist.ops.fp8 import quantize_mxfp8, gemm_mxfp8
from triton_dist.ops.gemm import GemmConfig
for (M, N, K) in [(4096, 12288, 6144), (8192, 8192, 8192)]:
    a = quantize_mxfp8(torch.randn(M, K, device="cuda", dtype=torch.bfloat16)); b = quantize_mxfp8(torch.randn(N, K, device="cuda", dtype=torch.bfloat16))
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for cg, bn in ((2, 128), (1, 256), (2, 256)):
        cfg = GemmConfig(bn=bn, cta_group=cg, group_m=8, use_tma_store=True)
        for _ in range(3): gemm_mxfp8(a, b, out=out, config=cfg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gemm_mxfp8(a, b, out=out, config=cfg)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps(dict(M=M, N=N, K=K, cg=cg, bn=bn, ms=ms, tflops=2 * M * N * K / ms / 1e9)))
PY
