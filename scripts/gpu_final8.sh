#!/bin/bash
# one 8-GPU call: all multi-rank correctness cases in a single torchrun, quick perf sweep, headline bench, Qwen3-8B TP8 decode
N=${1:-8}
mkdir -p gpurun_out
timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29611 \
    tests/dist_worker.py primitives allgather allreduce ag_gemm gemm_rs gemm_ar gemm_a2a moe ep_ll tp_e2e sp_pp ep_moe mega > gpurun_out/dist_all_n$N.log 2>&1
echo "dist rc=$?"; grep -E "CASE|Error|rank0\]:" gpurun_out/dist_all_n$N.log | grep -v "^W09" | tail -16 | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29612 scripts/gpu_sweep_dist.py quick > gpurun_out/sweep_n$N.log 2>&1
echo "sweep rc=$?"; grep -E "\"op\"" gpurun_out/sweep_n$N.log | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-2500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29614 scripts/bench_qwen3.py > gpurun_out/qwen3_n$N.log 2>&1
echo "qwen3 rc=$?"; tail -1 gpurun_out/qwen3_n$N.log | cut -c1-1500
