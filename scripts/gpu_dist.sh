#!/bin/bash
# multi-GPU checks: each case under its own timeout so a hang in one cannot eat the call
N=${1:-2}; shift
CASES=${@:-"primitives allgather allreduce ag_gemm gemm_rs"}
mkdir -p gpurun_out
for c in $CASES; do
  timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      tests/dist_worker.py $c > gpurun_out/dist_${c}_n$N.log 2>&1
  rc=$?
  echo "== case $c (N=$N) rc=$rc"
  grep -E "CASE|Error|error|rank0\]:|allreduce methods" gpurun_out/dist_${c}_n$N.log | grep -v "^W09" | tail -12 | cut -c1-300
done
