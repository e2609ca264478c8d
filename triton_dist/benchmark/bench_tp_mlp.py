"""TP MLP / attention block micro-benchmarks per forward mode (reference: benchmark/bench_tp_mlp.py, bench_tp_attn.py,
docs/getting-started/e2e/e2e_dense.md).  bash scripts/launch.sh --nproc_per_node=8 triton_dist/benchmark/bench_tp_mlp.py --M 2048"""
import argparse

import torch

import triton_dist.utils as U
from triton_dist.models import ARCHS
from triton_dist.parallel import TP_MLP
from triton_dist.profiler_utils import max_over_ranks, perf_func, print_benchmark_comparison


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Qwen/Qwen3-32B"); ap.add_argument("--M", type=int, default=2048); ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    U.initialize_distributed(heap_bytes=4 << 30)
    W, me = U.world_size(), U.rank()
    dev, grp = U.current_device(), U.get_triton_dist_world()
    a = ARCHS[args.model]
    mlp = TP_MLP(me, W, grp)
    g = torch.Generator(device=dev); g.manual_seed(me)
    mlp._init_parameters_from_shards((torch.randn(2 * a.intermediate_size // W, a.hidden_size, device=dev, generator=g) * 0.02).to(torch.bfloat16),
                                     (torch.randn(a.hidden_size, a.intermediate_size // W, device=dev, generator=g) * 0.02).to(torch.bfloat16))
    M = args.M
    mlp._init_ctx(M); mlp._init_AR_ctx(M); mlp._init_gemm_ar_ctx(M)
    x_full = torch.randn(M, a.hidden_size, device=dev, dtype=torch.bfloat16)
    x_shard = x_full[me * (M // W):(me + 1) * (M // W)].contiguous()
    row = {"name": f"{args.model} MLP M={M}"}
    for name, fn in (("torch", lambda: mlp.torch_fwd(x_full)), ("triton_dist", lambda: mlp.dist_triton_fwd(x_shard)),
                     ("triton_dist_AR", lambda: mlp.dist_triton_AR_fwd(x_full)), ("gemm_ar", lambda: mlp.dist_triton_gemm_ar_fwd(x_full))):
        _, t = perf_func(fn, args.iters, 5)
        row[name] = max_over_ranks(t, grp)
    if me == 0:
        print_benchmark_comparison([row], "torch", "TP MLP (ms, max over ranks)")
    mlp.finalize(); U.finalize_distributed()


if __name__ == "__main__":
    main()
