#!/bin/bash
# single-GPU confirmation round: all gpu tests, fp8 perf, ncu capture of the GEMM, Qwen3-8B decode (TP1, 12 layers)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --ignore=tests/test_flash_attn_gpu.py 2>&1 | tail -12
timeout 240 python -m pytest tests/test_flash_attn_gpu.py -m gpu -q -x -s 2>&1 | tail -25
bash scripts/gpu_fp8.sh 2>&1 | tail -8
timeout 300 python scripts/bench_qwen3.py --layers 12 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/gemm_prof python scripts/gpu_check_gemm.py 2 perf > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
