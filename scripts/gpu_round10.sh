#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29652 scripts/gpu_ag_multicast.py > gpurun_out/ag_multicast_n2.log 2>&1
echo "rc=$?"; grep -E "OK|Error|error|multicast|assert" gpurun_out/ag_multicast_n2.log | grep -v "^W09\|Warning" | tail -8 | cut -c1-400
