"""gemm_rs numerics + cost of the fp32 ring: error of the bf16 ring (W-1 roundings of the running partial), the fp32 ring (one
rounding at the owner) and NCCL's bf16 reduce_scatter, all against an fp32 golden (fp32 partial products reduced in fp32)."""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.ops.gemm_rs import create_gemm_rs_context, gemm_rs
U.initialize_distributed(seed=0, heap_bytes=3 << 30)
W, me = U.world_size(), U.rank()
dev, grp = U.current_device(), U.get_triton_dist_world()
M, N, K = 4096, 12288, 49152 // W
bf = torch.bfloat16
A = torch.randn(M, K, device=dev, dtype=bf) * 0.05
B = torch.randn(N, K, device=dev, dtype=bf) * 0.05
part = torch.zeros(M, N, device=dev, dtype=torch.float32)
kc = max(1, K // 4)
for k0 in range(0, K, kc):
    part.addmm_(A[:, k0:k0 + kc].float(), B[:, k0:k0 + kc].float().t())
gold = torch.empty(M // W, N, device=dev, dtype=torch.float32)
dist.reduce_scatter_tensor(gold, part, group=grp)
out_n = torch.empty(M // W, N, device=dev, dtype=bf)
dist.reduce_scatter_tensor(out_n, part.to(bf), group=grp)          # NCCL reducing bf16 partials
res = {}


def err(x):
    d = (x.float() - gold)
    t = torch.tensor([d.abs().max().item(), d.pow(2).mean().item()], device=dev)
    dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX, group=grp); dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM, group=grp)
    return round(t[0].item(), 5), round((t[1].item() / W) ** 0.5, 6)


def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(group=grp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
    return round(t.item(), 4)


res["nccl_bf16_reduce_scatter"] = dict(zip(("max_abs", "rms"), err(out_n)))
del part
for name, fp32 in (("ring_bf16", False), ("ring_fp32", True)):
    ctx = create_gemm_rs_context(M, N, output_dtype=bf, fp32_ring=fp32)
    o = gemm_rs(A, B.t(), ctx)
    torch.cuda.synchronize()
    res[name] = dict(zip(("max_abs", "rms"), err(o)))
    res[name]["ms"] = timed(lambda: gemm_rs(A, B.t(), ctx))
    U.barrier_all_host(); ctx.finalize()
res["gold_rms_magnitude"] = round(gold.pow(2).mean().sqrt().item(), 4)
if me == 0:
    print(json.dumps({"shape": f"gemm_rs M{M} N{N} K{49152} TP{W}", **res}))
    json.dump(res, open(f"gpurun_out/gemm_rs_numerics_n{W}.json", "w"), indent=1)
U.finalize_distributed()
