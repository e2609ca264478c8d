#!/bin/bash
# torchrun wrapper (reference: scripts/launch.sh:127-173).  Usage: bash scripts/launch.sh [--nproc_per_node=N] script.py args...
# Single node / single NVSwitch domain; rendezvous on 127.0.0.1.  TD_SYMM_HEAP_SIZE (default 4G) sizes the symmetric heap.
NPROC=${NPROC_PER_NODE:-$(nvidia-smi -L 2>/dev/null | wc -l)}
[ "$NPROC" -lt 1 ] && NPROC=2
ARGS=()
for a in "$@"; do
  case $a in
    --nproc_per_node=*|--nproc-per-node=*) NPROC="${a#*=}";;
    *) ARGS+=("$a");;
  esac
done
REPO_ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$REPO_ROOT${PYTHONPATH:+:$PYTHONPATH}"      # run from a checkout without installing the package
export TD_SYMM_HEAP_SIZE=${TD_SYMM_HEAP_SIZE:-${NVSHMEM_SYMMETRIC_SIZE:-4g}}
export CUDA_DEVICE_MAX_CONNECTIONS=${CUDA_DEVICE_MAX_CONNECTIONS:-1}
export NCCL_DEBUG=${NCCL_DEBUG:-ERROR}
PORT=${MASTER_PORT:-$((23000 + RANDOM % 2000))}
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node="$NPROC" --master-addr 127.0.0.1 --master-port "$PORT" "${ARGS[@]}"
