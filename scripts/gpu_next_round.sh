#!/bin/bash
# First GPU session of the next round: validate + time the opt-in candidates written after round 1's GPU budget was spent.
#   1 GPU :  bash scripts/gpu_next_round.sh 1     (192-wide GEMM tiles, flash v3 numerics + speed)
#   2 GPUs:  bash scripts/gpu_next_round.sh 2     (+ LL all-reduce numerics / graph latency)
#   8 GPUs:  bash scripts/gpu_next_round.sh 8     (+ multicast all-gather timing, bench with TD_AG_MULTICAST / TD_RS_BN192, isolated AG timings)
N=${1:-1}
mkdir -p gpurun_out
export TD_EXPERIMENTAL=1
timeout -k 10 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q -x -s 2>&1 | tail -25
if [ "$N" -ge 2 ]; then
  timeout -k 10 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29661 \
      tests/dist_worker.py allreduce_ll 2>&1 | grep -E "CASE|all_reduce 8 KB|Error" | tail -5
  timeout -k 10 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29662 \
      scripts/gpu_ag_multicast.py 2>&1 | grep -E "OK|multicast|Error" | tail -4
  TD_EP_NORMAL_GPU=1 timeout -k 10 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29663 \
      tests/dist_worker.py ep_moe 2>&1 | grep -E "CASE|Error" | tail -3
fi
if [ "$N" -ge 4 ]; then
  for knobs in "" "TD_AG_MULTICAST=1" "TD_RS_BN192=1"; do
    echo "== bench knobs: [$knobs]"
    env $knobs timeout -k 10 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29664 \
        bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-3000
  done
fi
