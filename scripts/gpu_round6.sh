#!/bin/bash
# 2-GPU round 6: AG-GEMM with direct local tiles + deep-unrolled push (numerics + bench), megakernel upgrades (1 GPU of the box)
mkdir -p gpurun_out
bash scripts/gpu_dist.sh 2 ag_gemm moe mega
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-1400
timeout -k 10 150 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "megakernel" 2>&1 | tail -4
timeout -k 10 150 python scripts/bench_qwen3.py --layers 12 2>&1 | tail -1
