"""Tutorial 08 -- write your own distributed kernel in CUDA C++ against the device header (csrc/td/primitives.cuh) and load it with
``triton_dist.jit.compile_cuda`` (the role of ``@triton_dist.jit`` / little_kernel's ``@ll_kernel`` in the reference).
    bash scripts/launch.sh --nproc_per_node=2 tutorials/08_custom_kernel.py        (GPUs; without one the kernel is only compiled)"""
import ctypes
import torch
import triton_dist.utils as U
from triton_dist.jit import compile_cuda, symm_ctx

SRC = r"""
#include "td/primitives.cuh"
using namespace td;
// every rank writes a tile into its successor's symmetric buffer, raises the successor's flag (release, system scope) and waits for
// its own flag (acquire) -- the notify / wait / symm_at pattern of tutorial 01, now inside one kernel
__global__ void ring(SymmCtx c, uint32_t* flag, float* data, uint32_t round) {
  const int nxt = (c.rank + 1) % c.world;
  symm_at(c, data, nxt)[threadIdx.x] = c.rank * 100.f + round;
  __syncthreads();
  if (threadIdx.x == 0) notify(c, flag, nxt, round);
  if (threadIdx.x < 32) wait<true, true>(flag, 1, round);       // flags carry the round number: nothing is ever reset
}
extern "C" void launch_ring(SymmCtx c, void* flag, void* data, unsigned round, void* stream) {
  ring<<<1, 64, 0, (cudaStream_t)stream>>>(c, (uint32_t*)flag, (float*)data, round);
}
"""

U.initialize_distributed()
W, me = U.world_size(), U.rank()
lib = compile_cuda(SRC, name="tutorial_ring")
if U.current_device().type != "cuda":
    U.dist_print("compiled for sm_100a; launching needs GPUs", allowed_ranks=[0])
else:
    data = U.nvshmem_create_tensor((64,), torch.float32)
    flag = U.nvshmem_create_tensor((1,), torch.int32)
    U.barrier_all_on_stream()
    lib.launch_ring.argtypes = [type(symm_ctx()), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
    for rnd in range(1, 6):
        lib.launch_ring(symm_ctx(), flag.data_ptr(), data.data_ptr(), rnd, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.all(data == ((me - 1) % W) * 100.0 + rnd), (rnd, data[:4])
        U.barrier_all_on_stream()               # the successor may overwrite my buffer only after I have checked it
    U.dist_print(f"rank {me}: 5 rounds through a user kernel OK")
U.finalize_distributed()
