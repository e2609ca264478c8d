#!/bin/bash
# final confirmation on 2 GPUs: every multi-rank case, bench --gpus 2, then the complete single-process gpu test suite
mkdir -p gpurun_out
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 \
    tests/dist_worker.py primitives allgather allreduce ag_gemm gemm_rs gemm_ar gemm_a2a moe moe_staged ep_ll tp_e2e sp_pp ep_moe mega > gpurun_out/dist_all_n2.log 2>&1
echo "dist rc=$?"; grep -E "CASE|Error|rank0\]:" gpurun_out/dist_all_n2.log | grep -v "^W09\|Warning" | tail -18 | cut -c1-200
timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-1300
CUDA_VISIBLE_DEVICES=0 timeout -k 10 300 python -m pytest tests -m gpu -q 2>&1 | tail -5
