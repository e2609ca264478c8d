#!/bin/bash
# standard single-GPU confirmation: gpu tests, smoke, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n1.json
