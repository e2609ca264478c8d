#!/bin/bash
# 8-GPU call A (most important first, tight timeouts): headline bench, Qwen3-8B TP8 decode, EP / GEMM+AR latencies + profiles
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout -k 10 170 bash -c "$(declare -f run); N=$N; run 29613 bench.py --gpus $N --steps 20 --warmup 5" > gpurun_out/bench_n$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-2600
timeout -k 10 200 bash -c "$(declare -f run); N=$N; run 29614 scripts/bench_qwen3.py" > gpurun_out/qwen3_n$N.log 2>&1
echo "qwen3 rc=$?"; tail -1 gpurun_out/qwen3_n$N.log | cut -c1-1500
timeout -k 10 110 bash -c "$(declare -f run); N=$N; run 29615 scripts/gpu_extras_dist.py" > gpurun_out/extras_n$N.log 2>&1
echo "extras rc=$?"; grep -E "\"op\"" gpurun_out/extras_n$N.log | cut -c1-600
