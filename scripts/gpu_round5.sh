#!/bin/bash
# single-GPU round 5: megakernel upgrades (fused norm, split-KV attention) -- numerics test + Qwen3-8B decode timing; flash variants
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "megakernel or engine or gdn or tma_gather" 2>&1 | tail -6
timeout -k 10 150 python scripts/bench_qwen3.py --layers 12 2>&1 | tail -1
