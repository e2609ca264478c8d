"""Tutorial 09 -- MoE at scale: the one-kernel ``moe_reduce_rs`` of a tensor-parallel MoE and Mega-EP, the expert-parallel MoE whose
dispatch runs inside the first grouped GEMM and whose combine runs inside the second (reference: kernels/nvidia/moe_reduce_rs.py,
ep_all2all_fused.py, function/nvidia/ep_moe_fused.py).
    bash scripts/launch.sh --nproc_per_node=2 tutorials/09_mega_ep_and_fused_moe.py"""
import torch
import torch.distributed as dist
import triton_dist.utils as U
from triton_dist.ops import ep_mega as EM
from triton_dist.ops import moe as M

U.initialize_distributed(seed=0)
W, me, dev = U.world_size(), U.rank(), U.current_device()
gpu = dev.type == "cuda"
dt = torch.bfloat16 if gpu else torch.float32
grp = U.get_triton_dist_world()

# ---- 1. tensor-parallel MoE down projection: grouped GEMM + weighted top-k reduce + reduce-scatter ------------------------------
# On a B200 this is ONE kernel (csrc/gemm_sm100.cuh, mode kMoeRS): tiles run column-chunk major, the epilogue adds every weighted
# row to its token's row of a symmetric partial with 16-byte L2 reductions, owners pull finished chunks through the NVSwitch.
T, H, I, E, topk = (2048 if gpu else 8) * W, (1024 if gpu else 16), (512 if gpu else 32), 8, 2
g = torch.Generator().manual_seed(1)                                    # the routing is identical on all ranks (TP)
ids = torch.rand(T, E, generator=g).topk(topk, dim=1).indices.to(torch.int32).to(dev)
wts = torch.softmax(torch.randn(T, topk, generator=g), -1).to(dev)
h = (torch.randn(T * topk, I // W, device=dev) * 0.5).to(dt)            # my K shard of the (token, k) activations
w_dn = (torch.randn(E, H, I // W, device=dev) * 0.2).to(dt)
rs = M.create_moe_rs_context(me, W, W, T * topk, H, E, topk, dt)
out = M.run_moe_reduce_rs(h, w_dn, ids, wts, rs)                        # [T / W, H]
gold = M.moe_reduce_rs_torch(h, w_dn.transpose(1, 2), ids, wts, grp, W, me)
U.dist_print(f"moe_reduce_rs: {tuple(out.shape)}, max err {(out.float() - gold.float()).abs().max().item():.3e}", allowed_ranks=[0])
U.barrier_all_host()
rs.finalize()

# ---- 2. Mega-EP: experts sharded over the ranks, tokens routed to them ------------------------------------------------------------
# The per-expert counts are all-gathered, so every sender knows the final expert-sorted row of each of its tokens on the
# destination: dispatch = stores straight into the grouped GEMM's A matrix (+ a return address); the down projection's epilogue
# stores every result row into its (token, k) slot on the owner.
T, H, I, epr, topk = (1024 if gpu else 12), (512 if gpu else 16), (256 if gpu else 8), 2, 2
E = epr * W
ctx = EM.create_ep_mega_context(T, H, topk, E, dt, capacity_factor=3.0)
gw = torch.Generator().manual_seed(2)
w_gu_all = (torch.randn(E, 2 * I, H, generator=gw) * 0.05).to(dt).to(dev)      # all experts (for the golden); I own a slice
w_dn_all = (torch.randn(E, H, I, generator=gw) * 0.05).to(dt).to(dev)
mine = slice(me * epr, (me + 1) * epr)
x = (torch.randn(T, H, device=dev) * 0.5).to(dt)
ids = torch.randn(T, E, device=dev).topk(topk, dim=1).indices.to(torch.int32)
wts = torch.softmax(torch.randn(T, topk, device=dev), -1)
h_sorted, handle = EM.mega_dispatch_group_gemm(ctx, x, ids, w_gu_all[mine].contiguous())          # half 1: dispatch || gate/up GEMM
from triton_dist.ops.elementwise import silu_mul
y = EM.mega_group_gemm_combine(ctx, silu_mul(h_sorted), handle, w_dn_all[mine].contiguous(), wts)   # half 2: down GEMM || combine
ref = EM.mega_ep_moe_reference(x, ids, wts, w_gu_all, w_dn_all)
U.dist_print(f"mega-EP: rows in my expert-sorted buffer {int(handle.n_rows)}, max err {(y.float() - ref).abs().max().item():.3e}")
U.barrier_all_host()
ctx.finalize()
U.dist_print("fused MoE tutorial OK", allowed_ranks=[0])
U.finalize_distributed()
