#!/bin/bash
# 8-GPU call C: headline bench with this round's AG-GEMM fixes, then the Qwen3-8B TP8 decode step (all backends + megakernel)
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout -k 10 150 bash -c "$(declare -f run); N=$N; run 29613 bench.py --gpus $N --steps 20 --warmup 5" > gpurun_out/bench_n$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-2600
timeout -k 10 130 bash -c "$(declare -f run); N=$N; run 29614 scripts/bench_qwen3.py" > gpurun_out/qwen3_n$N.log 2>&1
echo "qwen3 rc=$?"; tail -1 gpurun_out/qwen3_n$N.log | cut -c1-1500
