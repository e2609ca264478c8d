#!/bin/bash
# 2-GPU round 4: flipped defaults (fused MoE all-gather GEMM, single-kernel GEMM+AR in the TP layers, flash v2 prefill) end to end
bash scripts/gpu_dist.sh 2 moe moe_staged tp_e2e sp_pp
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29614 scripts/bench_qwen3.py --layers 12 > gpurun_out/qwen3_n2.log 2>&1
echo "qwen3 rc=$?"; tail -1 gpurun_out/qwen3_n2.log | cut -c1-1500
