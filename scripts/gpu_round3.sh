#!/bin/bash
# 2-GPU round: first GPU contact of the new multi-rank cases + the bench path with the transport autotune
bash scripts/gpu_dist.sh 2 gemm_ar gemm_a2a ag_gemm moe_fused sp_pp ep_moe mega tp_e2e
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-2500
timeout -k 10 200 python -m pytest tests/test_flash_attn_gpu.py -m gpu -q -x -s 2>&1 | tail -14
