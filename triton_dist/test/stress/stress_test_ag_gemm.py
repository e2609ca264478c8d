"""Stress test of the two flagship fused ops: random shapes on ONE pair of contexts, verification against NCCL / gloo + matmul, then many
unverified back-to-back calls that would expose a lost flag or a parity mix-up as a hang or as a wrong next result; optional straggler.

Reference: python/triton_dist/test/stress/stress_test_ag_gemm.py (ag_gemm only).  Here ``--op ag_gemm | gemm_rs | both``; runs on GPUs and on
the emulation backend (``TD_FORCE_HOST_BACKEND=1``), where ``TD_HOST_CHAOS_US`` adds random skew before every flag operation:

    bash scripts/launch.sh --nproc_per_node=8 triton_dist/test/stress/stress_test_ag_gemm.py --iters 100 --simulate_straggler
    TD_FORCE_HOST_BACKEND=1 TD_HOST_CHAOS_US=2000 bash scripts/launch.sh --nproc_per_node=3 triton_dist/test/stress/stress_test_ag_gemm.py \\
        --max_M 96 --N 48 --K 64 --iters 3 --verify_hang 4
"""
import argparse
import random
import sys

import torch
import torch.distributed as dist

import triton_dist.utils as U
from triton_dist.ops.ag_gemm import ag_gemm, create_ag_gemm_context
from triton_dist.ops.gemm_rs import create_gemm_rs_context, gemm_rs


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--op", default="both", choices=["ag_gemm", "gemm_rs", "both"])
    p.add_argument("--max_M", type=int, default=8192, help="largest global M (rounded down to the op's row granularity)")
    p.add_argument("--N", type=int, default=4096)
    p.add_argument("--K", type=int, default=4096)
    p.add_argument("--iters", type=int, default=20)
    p.add_argument("--verify_shapes", type=int, default=5)
    p.add_argument("--verify_hang", type=int, default=40)
    p.add_argument("--seed", type=int, default=40)
    p.add_argument("--simulate_straggler", action="store_true")
    return p.parse_args(argv)


def main(argv=None):
    a = parse_args(argv)
    U.initialize_distributed(seed=a.seed)
    W, me, dev = U.world_size(), U.rank(), U.current_device()
    gpu = dev.type == "cuda"
    grp = U.get_triton_dist_world()
    dtype = torch.bfloat16 if gpu else torch.float32
    gran = 128 * W if gpu else W                                   # rows: the fused kernels tile 128 rows per rank on hardware
    max_M = max(gran, a.max_M // gran * gran)
    N_loc, K_loc = max(8, a.N // W // 8 * 8), max(64, a.K // W // 64 * 64) if gpu else max(8, a.K // W)
    rng = random.Random(a.seed)                                     # same stream on every rank: shapes must agree
    ag_ctx = create_ag_gemm_context(max_M, N_loc, a.K, dtype) if a.op in ("ag_gemm", "both") else None
    rs_ctx = create_gemm_rs_context(max_M, a.N, me, W, W, dtype) if a.op in ("gemm_rs", "both") else None
    tol = dict(atol=0.5, rtol=3e-2) if gpu else dict(atol=1e-3, rtol=1e-3)
    g = torch.Generator().manual_seed(1000 + me)

    def data(M):
        x_ag = (torch.randn(M // W, a.K, generator=g) * 0.1 * (me + 1)).to(dtype).to(dev)
        w_ag = (torch.randn(N_loc, a.K, generator=g) * 0.1).to(dtype).to(dev)
        x_rs = (torch.randn(M, K_loc, generator=g) * 0.1).to(dtype).to(dev)
        w_rs = (torch.randn(a.N, K_loc, generator=g) * 0.1).to(dtype).to(dev)
        return x_ag, w_ag, x_rs, w_rs

    def check(M, x_ag, w_ag, x_rs, w_rs):
        if ag_ctx is not None:
            out = ag_gemm(x_ag, w_ag.t(), ag_ctx)
            full = torch.empty(M, a.K, dtype=dtype, device=dev)
            dist.all_gather_into_tensor(full, x_ag, group=grp)
            torch.testing.assert_close(out.float(), full.float() @ w_ag.float().t(), **tol)
        if rs_ctx is not None:
            out = gemm_rs(x_rs, w_rs.t(), rs_ctx)
            part = x_rs.float() @ w_rs.float().t()
            ref = torch.empty(M // W, a.N, dtype=torch.float32, device=dev)
            dist.reduce_scatter_tensor(ref, part, group=grp) if gpu else dist.reduce_scatter(ref, list(part.chunk(W)), group=grp)
            torch.testing.assert_close(out.float(), ref, **tol)

    for n in range(a.iters):
        for _ in range(a.verify_shapes):
            M = rng.randint(1, max_M // gran) * gran
            check(M, *data(M))
        straggler = (rng.randrange(W), rng.randint(10 ** 8, 2 * 10 ** 8)) if a.simulate_straggler else None
        M = rng.randint(1, max_M // gran) * gran
        x_ag, w_ag, x_rs, w_rs = data(M)
        for _ in range(a.verify_hang):                              # unverified back-to-back calls: flags / parity must survive
            if ag_ctx is not None:
                ag_gemm(x_ag, w_ag.t(), ag_ctx, straggler_option=straggler)
            if rs_ctx is not None:
                gemm_rs(x_rs, w_rs.t(), rs_ctx, straggler_option=straggler)
        check(M, x_ag, w_ag, x_rs, w_rs)                            # ... and the next verified call must still be right
        if gpu:
            torch.cuda.synchronize()
        U.dist_print(f"stress iteration {n + 1}/{a.iters} OK (last M = {M})", allowed_ranks=[0])
    U.barrier_all_on_stream()
    for c in (ag_ctx, rs_ctx):
        if c is not None:
            c.finalize()
    U.dist_print("stress test OK", allowed_ranks=[0])
    U.finalize_distributed()


if __name__ == "__main__":
    sys.exit(main())
