#!/bin/bash
# first contact: numerics for both CTA-group modes (each under its own timeout), then perf
mkdir -p gpurun_out
nvidia-smi -L | head -3
for cg in 1 2; do
  timeout 240 python scripts/gpu_check_gemm.py $cg > gpurun_out/gemm_cg$cg.log 2>&1; echo "cg$cg numerics exit=$?"
  tail -5 gpurun_out/gemm_cg$cg.log
done
for cg in 1 2; do
  timeout 300 python scripts/gpu_check_gemm.py $cg perf > gpurun_out/gemm_cg${cg}_perf.log 2>&1; echo "cg$cg perf exit=$?"
  tail -3 gpurun_out/gemm_cg${cg}_perf.log
done
