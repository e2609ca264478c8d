#!/bin/bash
# Race / memory checking recipe (reference: the commented compute-sanitizer line in scripts/launch.sh:160-162).
#   bash scripts/sanitize.sh memcheck  tests/test_gemm_gpu.py -k shapes     # single GPU
#   bash scripts/sanitize.sh racecheck tests/test_ops_gpu.py
#   bash scripts/sanitize.sh synccheck tests/test_ops_gpu.py
# Multi-rank protocols (flags, phases, rings) are additionally exercised without GPUs on the shared-memory emulation
# backend with wait timeouts (TD_HOST_TIMEOUT_US) that turn a hung flag wait into a TimeoutError instead of a spin.
TOOL=${1:-memcheck}; shift
exec compute-sanitizer --tool "$TOOL" --launch-timeout 120 python -m pytest -x -q -m gpu "$@"
