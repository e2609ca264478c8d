"""TP attention block micro-benchmark per forward mode, prefill and decode (reference: benchmark/bench_tp_attn.py,
docs/getting-started/e2e/e2e_dense.md:22-23,35-36).
bash scripts/launch.sh --nproc_per_node=8 triton_dist/benchmark/bench_tp_attn.py --bsz 32 --ctx 128 --mode prefill"""
import argparse

import torch

import triton_dist.utils as U
from triton_dist.models import ARCHS, KV_Cache
from triton_dist.parallel import TP_Attn
from triton_dist.profiler_utils import max_over_ranks, perf_func, print_benchmark_comparison


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Qwen/Qwen3-32B"); ap.add_argument("--bsz", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=128); ap.add_argument("--mode", default="prefill", choices=["prefill", "decode"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    U.initialize_distributed(heap_bytes=4 << 30)
    W, me = U.world_size(), U.rank()
    dev, grp = U.current_device(), U.get_triton_dist_world()
    a = ARCHS[args.model]
    Hq, Hkv, D, H = a.num_attention_heads, a.num_key_value_heads, a.head_dim, a.hidden_size
    attn = TP_Attn(me, W, grp)
    g = torch.Generator(device=dev); g.manual_seed(me)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    ones = torch.ones(D, device=dev, dtype=torch.bfloat16)
    attn._init_parameters_from_shards(rnd((Hq + 2 * Hkv) * D // W, H), rnd(H, Hq * D // W), ones if a.qk_norm else None,
                                      ones if a.qk_norm else None, Hq // W, max(1, Hkv // W), D, a.rms_norm_eps, a.rope_theta)
    bsz, q_len = args.bsz, (args.ctx if args.mode == "prefill" else 1)
    M = bsz * q_len
    attn._init_ctx(M); attn._init_AR_ctx(M); attn._init_gemm_ar_ctx(M)
    kv = KV_Cache(1, bsz, args.ctx + 8, Hkv, D, torch.bfloat16, W, dev)
    if args.mode == "decode":
        kv.rand_fill_kv_cache(args.ctx)
    x_full = torch.randn(bsz, q_len, H, device=dev, dtype=torch.bfloat16)
    x_shard = x_full[me * (bsz // W):(me + 1) * (bsz // W)].contiguous() if bsz % W == 0 else None
    pos = (kv.kv_offset.to(torch.int64)[:, None] + torch.arange(q_len, device=dev)[None]).contiguous()
    row = {"name": f"{args.model} attn {args.mode} bsz={bsz} ctx={args.ctx}"}
    modes = [("torch", lambda: attn.torch_fwd(x_full, pos, kv, 0)), ("triton_dist_AR", lambda: attn.dist_triton_AR_fwd(x_full, pos, kv, 0)),
             ("gemm_ar", lambda: attn.dist_triton_gemm_ar_fwd(x_full, pos, kv, 0))]
    if x_shard is not None:
        modes.insert(1, ("triton_dist", lambda: attn.dist_triton_fwd(x_shard, pos, kv, 0)))
    for name, fn in modes:
        _, t = perf_func(fn, args.iters, 5)
        row[name] = max_over_ranks(t, grp)
    if me == 0:
        print_benchmark_comparison([row], "torch", "TP attention (ms, max over ranks)")
    attn.finalize(); U.finalize_distributed()


if __name__ == "__main__":
    main()
