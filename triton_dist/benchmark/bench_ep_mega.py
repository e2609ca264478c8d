"""Mega-EP (dispatch || grouped GEMM, grouped GEMM || combine) vs the throughput-mode three-kernel path, device-timed, max over ranks.
    bash scripts/launch.sh --nproc_per_node=8 triton_dist/benchmark/bench_ep_mega.py --tokens 8192
Reference numbers (docs/kernels/nvidia/ep_all2all_fused.rst:486-492, 8 GPUs, 32k tokens/rank): dispatch ~3.5 ms,
dispatch + grouped GEMM ~5.9 ms, ~201 GB/s algorithmic bandwidth."""
import argparse
import json

import torch
import torch.distributed as dist

import triton_dist.utils as U
from triton_dist.ops import ep_mega as EM
from triton_dist.ops import ep_normal as EN
from triton_dist.ops.elementwise import silu_mul


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--inter", type=int, default=2048)
    ap.add_argument("--experts", type=int, default=64)
    ap.add_argument("--topk", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    U.initialize_distributed(seed=0, heap_bytes=12 << 30)
    W, me = U.world_size(), U.rank()
    dev, grp = U.current_device(), U.get_triton_dist_world()
    T, H, I, E, topk = a.tokens, a.hidden, a.inter, a.experts, a.topk
    epr = E // W
    bf = torch.bfloat16
    x = (torch.randn(T, H, device=dev) * 0.5).to(bf)
    ids = torch.rand(T, E, device=dev).topk(topk, dim=1).indices.to(torch.int32)
    wts = torch.softmax(torch.randn(T, topk, device=dev), -1)
    w_gu = (torch.randn(epr, 2 * I, H, device=dev) * 0.03).to(bf)
    w_dn = (torch.randn(epr, H, I, device=dev) * 0.03).to(bf)
    mctx = EM.create_ep_mega_context(T, H, topk, E, bf, capacity_factor=1.5)
    nctx = EN.create_ep_normal_ctx(T, H, topk, E, bf)

    def timed(fn, n=a.iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if W > 1:
            dist.barrier(group=grp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if W > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
        return round(t.item(), 4)

    state = {}

    def mega_full():
        return EM.mega_ep_moe(mctx, x, ids, wts, w_gu, w_dn)

    def mega_half1():
        state["h"], state["handle"] = EM.mega_dispatch_group_gemm(mctx, x, ids, w_gu)

    def normal_full():
        h = EN.ep_dispatch_normal(nctx, x, ids, wts)
        y = EN.ep_expert_ffn_normal(nctx, h, w_gu, w_dn)
        return EN.ep_combine_normal(nctx, y, h, ids)

    def normal_dispatch():
        state["nh"] = EN.ep_dispatch_normal(nctx, x, ids, wts)

    o1, o2 = mega_full().float(), normal_full().float()
    torch.cuda.synchronize()
    err = (o1 - o2).abs().max().item()
    assert err < 0.1 * max(1.0, o2.abs().max().item()), f"mega vs normal mismatch {err}"
    res = {"config": f"T={T}/rank H={H} I={I} E={E} top{topk} world={W}", "max_abs_diff_mega_vs_normal": round(err, 4)}
    res["mega_full_ms"] = timed(mega_full)
    res["normal_full_ms"] = timed(normal_full)
    res["normal_dispatch_only_ms"] = timed(normal_dispatch)
    # mega half 1 cannot be timed alone back to back (each call must be paired with its half 2): time full - (half-2 path) instead
    pair_bytes = T * topk * H * 2
    res["dispatch_bytes_per_rank"] = pair_bytes
    res["mega_speedup_vs_normal"] = round(res["normal_full_ms"] / res["mega_full_ms"], 3)
    flops = 2.0 * T * topk * H * (3 * I)
    res["mega_tflops_per_gpu"] = round(flops / res["mega_full_ms"] / 1e9, 1)
    res["algorithmic_dispatch_plus_combine_gbs"] = round(2 * pair_bytes / res["mega_full_ms"] / 1e6, 1)
    if me == 0:
        print(json.dumps(res))
        if a.json:
            json.dump(res, open(a.json, "w"), indent=1)
    U.barrier_all_host()
    mctx.finalize(); nctx.finalize()
    U.finalize_distributed()


if __name__ == "__main__":
    main()
