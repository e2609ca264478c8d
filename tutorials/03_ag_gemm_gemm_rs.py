"""Tutorial 03 -- AllGather-GEMM and GEMM-ReduceScatter (reference: tutorials/07,08; kernels/nvidia/allgather_gemm.py,
gemm_reduce_scatter.py).  The TP MLP pattern: ag_gemm -> activation -> gemm_rs, checked against NCCL + matmul."""
import torch
import torch.distributed as dist
import triton_dist.utils as U
from triton_dist.kernels.nvidia import ag_gemm, create_ag_gemm_context, create_gemm_rs_context, gemm_rs

U.initialize_distributed()
W, me = U.world_size(), U.rank()
dev = U.current_device()
big = dev.type == "cuda"
dt = torch.bfloat16 if big else torch.float32
M, H, I = (1024 * W, 1024, 2048) if big else (8 * W, 16, 32)
x = (torch.randn(M // W, H, device=dev) * 0.3).to(dt)
w_up = (torch.randn(I // W, H, device=dev) * 0.1).to(dt)       # column-parallel
w_dn = (torch.randn(H, I // W, device=dev) * 0.1).to(dt)       # row-parallel
ag_ctx = create_ag_gemm_context(M, I // W, H, dt)
rs_ctx = create_gemm_rs_context(M, H, output_dtype=dt)
h = ag_gemm(x, w_up.t(), ag_ctx)                               # [M, I/W]
y = gemm_rs(h, w_dn.t(), rs_ctx)                               # [M/W, H]
xf = torch.empty(M * H, dtype=dt, device=dev); dist.all_gather_into_tensor(xf, x.view(-1))
full = (xf.view(M, H).float() @ w_up.float().t()).to(dt).float() @ w_dn.float().t()
dist.all_reduce(full)
torch.testing.assert_close(y.float(), full[me * (M // W):(me + 1) * (M // W)], atol=0.5 if big else 1e-3, rtol=3e-2 if big else 1e-4)
U.dist_print("ag_gemm -> gemm_rs OK", allowed_ranks=[0])
ag_ctx.finalize(); rs_ctx.finalize(); U.finalize_distributed()
