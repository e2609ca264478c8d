#!/bin/bash
# first hardware contact of the throughput-mode EP kernels (2 GPUs, tight timeout)
mkdir -p gpurun_out
TD_EP_NORMAL_GPU=1 timeout -k 5 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29651 \
    tests/dist_worker.py ep_normal > gpurun_out/dist_ep_normal_n2.log 2>&1
echo "ep_normal rc=$?"; grep -E "CASE|Error|error|rank0\]:|ep_normal" gpurun_out/dist_ep_normal_n2.log | grep -v "^W09\|Warning" | tail -12 | cut -c1-300
