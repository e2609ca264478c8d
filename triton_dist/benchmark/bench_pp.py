"""Pipeline-parallel stage hand-off: symmetric-buffer put + signal vs NCCL send/recv (reference: benchmark/bench_pp.py,
layers/nvidia/pp_block.py:196-205).  Rank r sends to r+1 in a ring; time per hop, max over ranks.
bash scripts/launch.sh --nproc_per_node=8 triton_dist/benchmark/bench_pp.py --numel 4194304"""
import argparse

import torch

import triton_dist.utils as U
from triton_dist.parallel.pp import PPCommLayer
from triton_dist.profiler_utils import max_over_ranks, perf_func, print_benchmark_comparison


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, nargs="+", default=[8192, 1 << 20, 1 << 22, 1 << 24]); ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    U.initialize_distributed(heap_bytes=2 << 30)
    W, me = U.world_size(), U.rank()
    dev, grp = U.current_device(), U.get_triton_dist_world()
    rows = []
    for n in args.numel:
        x = torch.randn(n, device=dev, dtype=torch.bfloat16)
        row = {"name": f"{n * 2 / 2**20:.2f} MiB"}
        for backend in ("torch", "triton_dist"):
            pp = PPCommLayer(n, torch.bfloat16, me, W, backend=backend, group=grp)

            def hop():
                # even ranks send first, odd ranks receive first: a ring of blocking NCCL send/recv cannot deadlock
                if me % 2 == 0:
                    pp.send(x, (me + 1) % W); pp.recv((n,), torch.bfloat16, (me - 1) % W)
                else:
                    y = pp.recv((n,), torch.bfloat16, (me - 1) % W); pp.send(x, (me + 1) % W)
            if W > 1:
                _, t = perf_func(hop, args.iters, 3)
                row[backend] = max_over_ranks(t, grp)
            pp.finalize()
        rows.append(row)
    if me == 0 and W > 1:
        print_benchmark_comparison(rows, "torch", "PP hand-off per hop (ms, max over ranks)")
    U.finalize_distributed()


if __name__ == "__main__":
    main()
