#!/bin/bash
# single-GPU round 2: every gpu test on the CUDA backend (engine + megakernel included); first-contact kernels run in their
# own short-timeout processes so a hang cannot eat the call
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --ignore=tests/test_flash_attn_gpu.py -k "not tma_gather" 2>&1 | tail -15
timeout -k 10 240 python -m pytest tests/test_flash_attn_gpu.py -m gpu -q -x -s 2>&1 | tail -12
timeout -k 10 100 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "tma_gather" 2>&1 | tail -12
timeout -k 10 120 python scripts/bench_qwen3.py --layers 12 2>&1 | tail -1
