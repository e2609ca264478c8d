#!/bin/bash
export TD_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/bench_n2_final_quick.json 2> gpurun_out/bench_n2_final_quick.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_n2_final_quick.json
