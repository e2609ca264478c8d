"""Sweep AG-GEMM / GEMM-RS over the reference's LAYER_CONFIGS (reference: python/triton_dist/benchmark/
bench_allgather_gemm.py, test/utils.py:31-39) against NCCL + cuBLAS; csv on rank 0.
    bash scripts/launch.sh --nproc_per_node=8 triton_dist/benchmark/bench_allgather_gemm.py --M 8192 --csv out.csv"""
import argparse
import csv

import torch
import torch.distributed as dist

import triton_dist.utils as U
from triton_dist.kernels.nvidia import ag_gemm, create_ag_gemm_context, create_gemm_rs_context, gemm_rs
from triton_dist.profiler_utils import max_over_ranks, perf_func
from triton_dist.test.utils import LAYER_CONFIGS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=8192); ap.add_argument("--iters", type=int, default=10); ap.add_argument("--csv", default="")
    args = ap.parse_args()
    U.initialize_distributed(heap_bytes=8 << 30)
    W, me = U.world_size(), U.rank()
    dev, grp = U.current_device(), U.get_triton_dist_world()
    rows = []
    for name, c in LAYER_CONFIGS.items():
        N, K, M = c["N"], c["K"], args.M
        if (N // W) % 8 or (K // W) % 8 or (M // W) % 128:
            continue
        # AG-GEMM: A [M/W, K] x B [N/W, K]
        A = torch.randn(M // W, K, device=dev, dtype=torch.bfloat16) * 0.05
        B = torch.randn(N // W, K, device=dev, dtype=torch.bfloat16) * 0.05
        ctx = create_ag_gemm_context(M, N // W, K, torch.bfloat16)
        full = torch.empty(M, K, device=dev, dtype=torch.bfloat16)

        def nccl_ag():
            dist.all_gather_into_tensor(full, A, group=grp)
            return torch.matmul(full, B.t())
        _, t_ref = perf_func(nccl_ag, args.iters, 3)
        _, t_our = perf_func(lambda: ag_gemm(A, B.t(), ctx), args.iters, 3)
        t_ref, t_our = max_over_ranks(t_ref, grp), max_over_ranks(t_our, grp)
        rows.append(dict(op="ag_gemm", model=name, M=M, N=N, K=K, nccl_cublas_ms=t_ref, ours_ms=t_our, speedup=t_ref / t_our,
                         tflops=2.0 * M * N * K / t_our / 1e9))
        U.barrier_all_host(); ctx.finalize()
        # GEMM-RS: A [M, N/W] x B [K, N/W]  (the down projection: reduce over the sharded N)
        A2 = torch.randn(M, N // W, device=dev, dtype=torch.bfloat16) * 0.05
        B2 = torch.randn(K, N // W, device=dev, dtype=torch.bfloat16) * 0.05
        rs = create_gemm_rs_context(M, K, output_dtype=torch.bfloat16)
        out = torch.empty(M // W, K, device=dev, dtype=torch.bfloat16)

        def nccl_rs():
            dist.reduce_scatter_tensor(out, torch.matmul(A2, B2.t()), group=grp)
        _, t_ref = perf_func(nccl_rs, args.iters, 3)
        _, t_our = perf_func(lambda: gemm_rs(A2, B2.t(), rs), args.iters, 3)
        t_ref, t_our = max_over_ranks(t_ref, grp), max_over_ranks(t_our, grp)
        rows.append(dict(op="gemm_rs", model=name, M=M, N=K, K=N, nccl_cublas_ms=t_ref, ours_ms=t_our, speedup=t_ref / t_our,
                         tflops=2.0 * M * N * K / t_our / 1e9))
        U.barrier_all_host(); rs.finalize()
    if me == 0:
        for r in rows:
            print(r)
        if args.csv:
            with open(args.csv, "w", newline="") as f:
                w = csv.DictWriter(f, fieldnames=list(rows[0]))
                w.writeheader(); w.writerows(rows)
    U.finalize_distributed()


if __name__ == "__main__":
    main()
