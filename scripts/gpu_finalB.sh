#!/bin/bash
# N-GPU call B: every multi-rank correctness case in one torchrun
N=${1:-4}
mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29611 \
    tests/dist_worker.py primitives allgather allreduce ag_gemm gemm_rs gemm_ar gemm_a2a moe moe_staged ep_ll tp_e2e sp_pp ep_moe mega > gpurun_out/dist_all_n$N.log 2>&1
echo "dist rc=$?"; grep -E "CASE|Error|rank0\]:" gpurun_out/dist_all_n$N.log | grep -v "^W09" | tail -18 | cut -c1-250
