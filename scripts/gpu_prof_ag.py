"""Intra-kernel profile of one fused ag_gemm call (per-stream push / completion times, flag waits, tiles)."""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.ops.ag_gemm import ag_gemm, create_ag_gemm_context
from triton_dist.ops.gemm import GemmConfig
from triton_dist.tools.profiler import ProfilerBuffer, export_to_perfetto_trace, summarize
U.initialize_distributed(seed=0)
W, me = U.world_size(), U.rank()
M, N, K = 4096, 4096, 4096
ctx = create_ag_gemm_context(M, N // W, K, torch.bfloat16)
A = torch.randn(M // W, K, device="cuda", dtype=torch.bfloat16); B = torch.randn(N // W, K, device="cuda", dtype=torch.bfloat16)
cfg = GemmConfig(bn=256, cta_group=2, group_m=max(1, (M // W) // 256), n_comm_ctas=16)
for _ in range(3): ag_gemm(A, B.t(), ctx, gemm_config=cfg)
pb = ProfilerBuffer()
torch.cuda.synchronize(); dist.barrier()
ag_gemm(A, B.t(), ctx, gemm_config=cfg, profiler=pb)
torch.cuda.synchronize()
if me == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    print(json.dumps(summarize(pb), indent=1))
    export_to_perfetto_trace(pb, f"gpurun_out/ag_gemm_trace_n{W}.json.gz", rank=me)
    ev = pb.events(); t0 = min(e["ns"] for e in ev); t1 = max(e["ns"] for e in ev)
    print("kernel span us:", (t1 - t0) / 1e3)
    last_push = max(e["ns"] for e in ev if e["tag"] == 2 and not e["start"])
    print("last push completion at us:", (last_push - t0) / 1e3)
U.finalize_distributed()
