"""First contact of the NVLS multicast all-gather transport of ag_gemm: numerics vs NCCL all-gather + matmul, then device time vs the SM push."""
import json, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.ops.ag_gemm import ag_gemm, create_ag_gemm_context, default_ag_config
from triton_dist.ops.gemm import GemmConfig
U.initialize_distributed(seed=0, heap_bytes=2 << 30)
W, me = U.world_size(), U.rank()
dev, grp = U.current_device(), U.get_triton_dist_world()
bf = torch.bfloat16
for (M, N, K) in ((512 * W, 512, 1024), (256 * W, 256, 512), (4096, 4096 // W, 4096)):
    ctx = create_ag_gemm_context(M, N, K, bf)
    ctx.workspace.view(torch.int16).fill_(0x7FC0)
    for it in range(4):
        A = (torch.randn(M // W, K, device=dev) * 0.5).to(bf); Wt = (torch.randn(N, K, device=dev) * 0.5).to(bf)
        if it == 2 and me == 1:
            torch.cuda._sleep(2_000_000)
        C = ag_gemm(A, Wt.t(), ctx, transport="multicast")
        full = torch.empty(M * K, device=dev, dtype=bf)
        dist.all_gather_into_tensor(full, A.view(-1), group=grp)
        ref = full.view(M, K).float() @ Wt.float().t()
        err = (C.float() - ref).abs().max().item()
        assert err < 0.5 + 2e-2 * ref.abs().max().item(), (M, N, K, it, err)
    U.barrier_all_host()
    if (M, K) == (4096, 4096):
        def timed(fn, n=20):
            for _ in range(5): fn()
            torch.cuda.synchronize(); dist.barrier(group=grp); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
            return t.item() * 1e3
        base = default_ag_config(M, N, K, W)
        res = {"W": W, "sm_us": timed(lambda: ag_gemm(A, Wt.t(), ctx))}
        for nc in (4, 8, 16):
            cfg = GemmConfig(base.bn, base.cta_group, base.group_m, True, 0, nc)
            res[f"multicast_nc{nc}_us"] = timed(lambda: ag_gemm(A, Wt.t(), ctx, gemm_config=cfg, transport="multicast"))
        if me == 0: print(json.dumps(res))
    ctx.finalize()
if me == 0: print("multicast transport OK")
U.finalize_distributed()
