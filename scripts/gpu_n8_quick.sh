#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
bash scripts/gpu_dist.sh $N ag_gemm
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29801 scripts/gpu_sweep_dist.py ${2:-quick} > gpurun_out/sweep_n$N.log 2>&1; echo "sweep rc=$?"; grep -v "^W09\|^\[W" gpurun_out/sweep_n$N.log | grep -E "ag_gemm|gemm_rs|Error" | cut -c1-330
