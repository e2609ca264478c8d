"""Tutorial 04 -- TP-MoE (ag_group_gemm + moe_reduce_rs) and expert-parallel low-latency dispatch/combine
(reference: tutorials/09,10 + kernels/nvidia/allgather_group_gemm.py, moe_reduce_rs.py, low_latency_all_to_all_v2.py)."""
import torch
import triton_dist.utils as U
from triton_dist.parallel import EP_MoE

U.initialize_distributed()
W, me = U.world_size(), U.rank()
dev = U.current_device()
dt = torch.bfloat16 if dev.type == "cuda" else torch.float32
E, H, I, topk, T = 4 * W, 256, 128, 2, 32
g = torch.Generator().manual_seed(0)
router = (torch.randn(E, H, generator=g) * 0.5).to(dt).to(dev)
gu = (torch.randn(E, 2 * I, H, generator=g) * 0.1).to(dt).to(dev)
dn = (torch.randn(E, H, I, generator=g) * 0.1).to(dt).to(dev)
epr = E // W
moe = EP_MoE(me, W, U.get_triton_dist_world())
moe._init_parameters_from_shards(router, gu[me * epr:(me + 1) * epr].contiguous(), dn[me * epr:(me + 1) * epr].contiguous(), topk)
moe._init_ctx(T)
x = (torch.randn(T, H, generator=torch.Generator().manual_seed(1 + me)) * 0.5).to(dt).to(dev)
torch.testing.assert_close(moe.dist_triton_fwd(x).float(), moe.torch_fwd(x).float(), atol=5e-2, rtol=5e-2)
U.dist_print("EP MoE (dispatch -> grouped GEMM -> combine) OK", allowed_ranks=[0])
moe.finalize(); U.finalize_distributed()
