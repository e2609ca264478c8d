"""Intra-kernel profile of one fused ag_gemm call (comm-CTA issue / fence+publish times, producer flag waits, tiles).
    torchrun ... scripts/gpu_prof_ag.py [transport=multicast] [bn=128] [cta_group=1] [n_comm=16] [kslices=8]"""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.ops.ag_gemm import ag_gemm, create_ag_gemm_context
from triton_dist.ops.gemm import GemmConfig
from triton_dist.tools.profiler import ProfilerBuffer, export_to_perfetto_trace, summarize
kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
transport, bn, cg = kv.get("transport", "multicast"), int(kv.get("bn", 128)), int(kv.get("cta_group", 1))
n_comm, ks, gr = int(kv.get("n_comm", 24)), int(kv.get("kslices", 8)), int(kv.get("groups", 3))
tail = int(kv.get("tail", 0))
U.initialize_distributed(seed=0)
W, me = U.world_size(), U.rank()
M, N, K = 4096, 4096, 4096
ctx = create_ag_gemm_context(M, N // W, K, torch.bfloat16)
A = torch.randn(M // W, K, device="cuda", dtype=torch.bfloat16); B = torch.randn(N // W, K, device="cuda", dtype=torch.bfloat16)
cfg = GemmConfig(bn=bn, cta_group=cg, group_m=max(1, (M // W) // (128 * cg)), n_comm_ctas=n_comm)
pb = ProfilerBuffer()
for _ in range(3): ag_gemm(A, B.t(), ctx, gemm_config=cfg, transport=transport, kslices=ks, comm_groups=gr, tail_pct=tail)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
# warm calls right before the profiled one (an isolated call after an idle gap runs its first ~100 us at ramping clocks)
for _ in range(6): ag_gemm(A, B.t(), ctx, gemm_config=cfg, transport=transport, kslices=ks, comm_groups=gr, tail_pct=tail)
ag_gemm(A, B.t(), ctx, gemm_config=cfg, transport=transport, kslices=ks, comm_groups=gr, tail_pct=tail, profiler=pb)
torch.cuda.synchronize()
if me == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    print(json.dumps(summarize(pb), indent=1))
    tag = f"{transport}_bn{bn}_cg{cg}_nc{n_comm}_ks{ks}_g{gr}_t{tail}"
    export_to_perfetto_trace(pb, f"gpurun_out/ag_gemm_trace_n{W}_{tag}.json.gz", rank=me)
    ev = pb.events(); t0 = min(e["ns"] for e in ev); t1 = max(e["ns"] for e in ev)
    print("kernel span us:", (t1 - t0) / 1e3)
    for tg, name in ((1, "push issue"), (6, "fence+publish"), (3, "producer flag wait"), (4, "mainloop"), (5, "epilogue")):
        ends = [e["ns"] for e in ev if e["tag"] == tg and not e["start"]]
        starts = [e["ns"] for e in ev if e["tag"] == tg and e["start"]]
        if ends:
            print(f"{name}: n={len(ends)} first start {(min(starts) - t0) / 1e3:.1f} us, last end {(max(ends) - t0) / 1e3:.1f} us")
U.finalize_distributed()
