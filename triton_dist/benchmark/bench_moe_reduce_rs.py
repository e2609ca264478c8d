"""MoE down projection + top-k reduce + ReduceScatter (BASELINE config #4: Mixtral-8x7B, TP = world size):
T tokens, top-2 of 8 experts, hidden N = 4096, intermediate 14336 sharded over the ranks (K = 14336 / W).

    bash scripts/launch.sh --nproc_per_node=8 triton_dist/benchmark/bench_moe_reduce_rs.py --tokens 8192

Three implementations, device-timed (CUDA events, rotating inputs, max over ranks), all checked against the masked-matmul
golden (reference: test/nvidia/test_moe_reduce_rs.py:88-107):
  * fused   : ONE kernel (csrc/gemm_sm100.cuh mode kMoeRS; reference moe_reduce_rs.py:168-246 + 549-619 needs two kernels + streams)
  * staged  : grouped GEMM kernel -> reduce_topk kernel -> NVLS reduce-scatter kernel per N chunk on a side stream (round-1 path)
  * nccl    : grouped GEMM kernel -> reduce_topk kernel -> NCCL reduce_scatter_tensor
"""
import argparse
import json
import os

import torch
import torch.distributed as dist

import triton_dist.utils as U
from triton_dist.ops import moe as M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--inter", type=int, default=14336)
    ap.add_argument("--experts", type=int, default=8)
    ap.add_argument("--topk", type=int, default=2)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    U.initialize_distributed(seed=0, heap_bytes=4 << 30)
    W, me = U.world_size(), U.rank()
    dev, grp = U.current_device(), U.get_triton_dist_world()
    T, H, E, topk = args.tokens, args.hidden, args.experts, args.topk
    K = args.inter // W
    bf = torch.bfloat16
    nset = 3
    g = torch.Generator(device="cpu").manual_seed(1)
    ids = torch.rand(T, E, generator=g).topk(topk, dim=1).indices.to(torch.int32).to(dev)      # same routing on all ranks
    wts = torch.softmax(torch.randn(T, topk, generator=g), -1).to(dev)
    xs = [(torch.randn(T * topk, K, device=dev) * 0.3).to(bf) for _ in range(nset)]
    ws = [(torch.randn(E, H, K, device=dev) * 0.05).to(bf) for _ in range(nset)]
    ctx = M.create_moe_rs_context(me, W, W, T * topk, H, E, topk, bf)

    def fused(i):
        return M.run_moe_reduce_rs(xs[i % nset], ws[i % nset], ids, wts, ctx)

    def staged(i):
        os.environ["TD_MOE_RS_FUSED"] = "0"
        try:
            return M.run_moe_reduce_rs(xs[i % nset], ws[i % nset], ids, wts, ctx, n_chunks=2)
        finally:
            os.environ["TD_MOE_RS_FUSED"] = "1"

    out_n = torch.empty(T // W, H, device=dev, dtype=bf)

    def nccl(i):
        part = M._moe_down_partial(xs[i % nset], ws[i % nset], ids, wts, ctx)
        if W > 1:
            dist.reduce_scatter_tensor(out_n, part, group=grp)
            return out_n
        return part

    gold = M.moe_reduce_rs_torch(xs[0], ws[0].transpose(1, 2), ids, wts, grp, W, me)
    errs = {}
    for name, fn in (("fused", fused), ("staged", staged), ("nccl", nccl)):
        o = fn(0)
        torch.cuda.synchronize()
        errs[name] = (o.float() - gold).abs().max().item()
        assert torch.allclose(o.float(), gold, atol=0.25, rtol=3e-2), f"{name}: max abs err {errs[name]}"

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        if W > 1:
            dist.barrier(group=grp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            fn(3 + i)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        if W > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
        return t.item()

    res = {"config": f"Mixtral-8x7B down-proj + topk reduce + RS: T={T} top{topk}/{E} N={H} K={args.inter}/{W}", "world": W}
    flops = 2.0 * T * topk * H * K
    sweep = {}
    for ncomm in (8, 16, 24, 32):          # comm CTAs of the fused kernel (the rest of the SMs run GEMM tiles)
        ctx.n_comm = ncomm
        sweep[ncomm] = round(timed(fused), 4)
    ctx.n_comm = min(sweep, key=sweep.get)
    for name, fn in (("fused", fused), ("staged", staged), ("nccl", nccl)):
        ms = timed(fn)
        res[name] = {"ms": round(ms, 4), "tflops_per_gpu": round(flops / ms / 1e9, 1), "max_abs_err": round(errs[name], 4)}
    res["fused"]["n_comm_ctas"] = ctx.n_comm
    res["fused"]["ms_by_n_comm_ctas"] = sweep
    res["fused_vs_staged"] = round(res["staged"]["ms"] / res["fused"]["ms"], 3)
    res["fused_vs_nccl"] = round(res["nccl"]["ms"] / res["fused"]["ms"], 3)
    # roofline: grouped GEMM at the measured bf16 peak vs the bytes every rank must ship ((W-1)/W of its [T, N] partial)
    link, peak = 770e9, 1676.7e12
    res["roofline_ms"] = round(max(flops / peak, (W - 1) / W * T * H * 2 / link) * 1e3, 4)
    res["fused_frac_of_roofline"] = round(res["roofline_ms"] / res["fused"]["ms"], 3)
    if me == 0:
        print(json.dumps(res))
        if args.json:
            json.dump(res, open(args.json, "w"), indent=1)
    U.barrier_all_host()
    ctx.finalize()
    U.finalize_distributed()


if __name__ == "__main__":
    main()
