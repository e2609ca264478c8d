#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
bash scripts/gpu_dist.sh $N ag_gemm gemm_rs allreduce
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29801 scripts/gpu_sweep_dist.py $2 > gpurun_out/sweep_n$N.log 2>&1; echo "sweep rc=$?"; grep -v "^W09\|^\[W" gpurun_out/sweep_n$N.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log
