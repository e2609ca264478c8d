#!/bin/bash
# single-GPU round 2: every gpu test on the CUDA backend (engine + megakernel included), flash-attention both tile widths
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --ignore=tests/test_flash_attn_gpu.py 2>&1 | tail -15
timeout 300 python -m pytest tests/test_flash_attn_gpu.py -m gpu -q -x -s 2>&1 | tail -12
