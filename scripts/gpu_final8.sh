#!/bin/bash
# one N-GPU call, most important first, every step under a tight timeout (a hang must not eat the GPU budget):
# headline bench -> multi-rank correctness -> perf sweep -> Qwen3-8B TP decode -> EP / GEMM+AR latencies
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout -k 10 200 bash -c "$(declare -f run); N=$N; run 29613 bench.py --gpus $N --steps 20 --warmup 5" > gpurun_out/bench_n$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-2600
timeout -k 10 330 bash -c "$(declare -f run); N=$N; run 29611 tests/dist_worker.py primitives allgather allreduce ag_gemm gemm_rs gemm_ar gemm_a2a moe moe_staged ep_ll tp_e2e sp_pp ep_moe mega" > gpurun_out/dist_all_n$N.log 2>&1
echo "dist rc=$?"; grep -E "CASE|Error|rank0\]:" gpurun_out/dist_all_n$N.log | grep -v "^W09" | tail -18 | cut -c1-250
timeout -k 10 170 bash -c "$(declare -f run); N=$N; run 29612 scripts/gpu_sweep_dist.py quick" > gpurun_out/sweep_n$N.log 2>&1
echo "sweep rc=$?"; grep -E "\"op\"" gpurun_out/sweep_n$N.log | cut -c1-300
timeout -k 10 240 bash -c "$(declare -f run); N=$N; run 29614 scripts/bench_qwen3.py" > gpurun_out/qwen3_n$N.log 2>&1
echo "qwen3 rc=$?"; tail -1 gpurun_out/qwen3_n$N.log | cut -c1-1500
timeout -k 10 110 bash -c "$(declare -f run); N=$N; run 29615 scripts/gpu_extras_dist.py" > gpurun_out/extras_n$N.log 2>&1
echo "extras rc=$?"; grep -E "\"op\"" gpurun_out/extras_n$N.log | cut -c1-400
