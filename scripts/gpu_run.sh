#!/bin/bash
# GPU validation stages (results land in gpurun_out/; the summaries that matter are copied to profiles/).
#   gpurun            -- bash scripts/gpu_run.sh a        1 GPU : smoke, GEMM tests, bench (ours + reference arm), all GPU tests
#   gpurun --gpus N   -- bash scripts/gpu_run.sh b|d|e|f|g N    : distributed cases / op benchmarks / traces at N = 2
#   gpurun --gpus 8   -- bash scripts/gpu_run.sh c|h 8          : full bench, all distributed cases, MoE / Mega-EP / Qwen3 benchmarks
# every multi-rank command runs under its own `timeout` (a hang must never reach gpurun's limit)
set -x
mkdir -p gpurun_out
export TD_NO_AUTOBUILD=1
export PYTHONPATH=$PWD:$PYTHONPATH
stage=${1:-a}
if [ "$stage" = "a" ]; then
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
  timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/test_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -5 gpurun_out/test_gemm.log
  timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref_n1.json
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json
  TD_SPLITK=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --quick > gpurun_out/bench_n1_nosplit.json 2> gpurun_out/bench_n1_nosplit.err; tail -c 600 gpurun_out/bench_n1_nosplit.json
  timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gemm_gpu.py > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -8 gpurun_out/test_gpu.log
fi
if [ "$stage" = "b" ]; then
  N=${2:-2}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py ag_gemm gemm_rs moe_rs moe > gpurun_out/dist_ag_rs_n$N.log 2>&1; echo "dist rc=$?"; tail -15 gpurun_out/dist_ag_rs_n$N.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
fi
if [ "$stage" = "c" ]; then
  N=${2:-8}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 600 $TR --master-port 29521 tests/dist_worker.py ag_gemm gemm_rs moe_rs gemm_ar moe ep_ll ep_normal mega tp_e2e allreduce > gpurun_out/dist_n$N.log 2>&1; echo "dist rc=$?"; tail -12 gpurun_out/dist_n$N.log
  timeout 500 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
  timeout 300 $TR --master-port 29523 triton_dist/benchmark/bench_moe_reduce_rs.py --json gpurun_out/moe_reduce_rs_n$N.json > gpurun_out/moe_rs_n$N.log 2>&1; echo "moe rc=$?"; tail -3 gpurun_out/moe_rs_n$N.log
  timeout 200 $TR --master-port 29524 scripts/gpu_prof_ag.py transport=sm_k bn=256 cta_group=2 n_comm=32 kslices=2 groups=1 > gpurun_out/prof_ag_mc_n$N.log 2>&1; echo "prof rc=$?"; tail -8 gpurun_out/prof_ag_mc_n$N.log
  timeout 200 $TR --master-port 29526 scripts/gpu_prof_ag.py transport=sm_k bn=256 cta_group=2 n_comm=48 kslices=4 groups=2 > gpurun_out/prof_ag_mc2_n$N.log 2>&1; echo "prof2 rc=$?"; tail -8 gpurun_out/prof_ag_mc2_n$N.log
fi
if [ "$stage" = "d" ]; then
  N=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 400 $TR --master-port 29531 tests/dist_worker.py ag_gemm moe tp_e2e > gpurun_out/dist_d_n$N.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/dist_d_n$N.log
  timeout 300 $TR --master-port 29533 triton_dist/benchmark/bench_moe_reduce_rs.py --json gpurun_out/moe_reduce_rs_n$N.json > gpurun_out/moe_rs_n$N.log 2>&1; echo "moe rc=$?"; tail -3 gpurun_out/moe_rs_n$N.log
  timeout 200 $TR --master-port 29534 scripts/gpu_prof_ag.py transport=sm_k bn=256 cta_group=2 n_comm=32 kslices=2 groups=1 > gpurun_out/prof_ag_mc_n$N.log 2>&1; echo "prof rc=$?"; tail -8 gpurun_out/prof_ag_mc_n$N.log
fi
if [ "$stage" = "e" ]; then
  N=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 500 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --quick > gpurun_out/bench_quick_n$N.json 2> gpurun_out/bench_quick_n$N.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_quick_n$N.json; tail -3 gpurun_out/bench_quick_n$N.err
  for cfg in "transport=sm_k bn=256 cta_group=2 n_comm=32 kslices=2 groups=1" "transport=sm_k bn=256 cta_group=2 n_comm=32 kslices=4 groups=2" "transport=sm bn=256 cta_group=2 n_comm=32"; do
    timeout 200 $TR --master-port 29544 scripts/gpu_prof_ag.py $cfg > gpurun_out/prof_e.log 2>&1; echo "prof [$cfg] rc=$?"; tail -6 gpurun_out/prof_e.log
  done
fi
if [ "$stage" = "f" ]; then
  N=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 90 $TR --master-port 29551 tests/dist_worker.py ep_mega > gpurun_out/dist_f_n$N.log 2>&1; echo "dist rc=$?"; grep -v "^\[W\|^W0\|OMP_NUM\|^\*\*" gpurun_out/dist_f_n$N.log | tail -25
fi
if [ "$stage" = "g" ]; then
  N=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "decode or engine or mega" > gpurun_out/test_decode.log 2>&1; echo "decode rc=$?"; tail -4 gpurun_out/test_decode.log
  for c in ep_mega gemm_q8 sp_pp; do
    timeout 150 $TR --master-port 29561 tests/dist_worker.py $c > gpurun_out/dist_g_$c.log 2>&1; echo "dist $c rc=$?"; grep -v "^\[W\|^W0\|OMP_NUM\|^\*\*" gpurun_out/dist_g_$c.log | grep "CASE\|Error\|error" | head -8
  done
fi
if [ "$stage" = "h" ]; then
  N=${2:-8}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 420 $TR --master-port 29571 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
  for c in "ag_gemm gemm_rs moe_rs gemm_ar gemm_q8" "moe ep_ll ep_normal ep_mega" "mega tp_e2e sp_pp allreduce"; do
    timeout 240 $TR --master-port 29572 tests/dist_worker.py $c > gpurun_out/dist_h_tmp.log 2>&1; echo "dist [$c] rc=$?"; cat gpurun_out/dist_h_tmp.log >> gpurun_out/dist_all_n$N.log; grep "CASE\|Error" gpurun_out/dist_h_tmp.log | head -12
  done
  timeout 200 $TR --master-port 29573 triton_dist/benchmark/bench_moe_reduce_rs.py --json gpurun_out/moe_reduce_rs_n$N.json > gpurun_out/moe_rs_n$N.log 2>&1; echo "moe rc=$?"; tail -2 gpurun_out/moe_rs_n$N.log | cut -c1-900
  timeout 240 $TR --master-port 29574 triton_dist/benchmark/bench_ep_mega.py --json gpurun_out/ep_mega_n$N.json > gpurun_out/ep_mega_n$N.log 2>&1; echo "ep_mega rc=$?"; tail -2 gpurun_out/ep_mega_n$N.log | cut -c1-900
  timeout 240 $TR --master-port 29575 scripts/bench_qwen3.py > gpurun_out/qwen3_n$N.log 2>&1; echo "qwen3 rc=$?"; tail -4 gpurun_out/qwen3_n$N.log | cut -c1-900
fi
