#!/bin/bash
# 2-GPU sanity: fused GEMM+AR after the comm-side change, smoke(), device microbenchmarks
mkdir -p gpurun_out
bash scripts/gpu_dist.sh 2 gemm_ar tp_e2e
timeout -k 10 120 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-400
timeout -k 10 120 python -m triton_dist.benchmark.microbench 2>&1 | tail -1 | tee gpurun_out/microbench.json
