"""Offline GEMM config search (reference: python/triton_dist/tools/tune/tune_gemm.py, find_topk.py).
    python -m triton_dist.tools.tune.tune_gemm --shapes 4096x4096x4096 8192x8192x8192 --topk 3"""
import argparse
import itertools
import json

import torch

from triton_dist.ops.gemm import GemmConfig, gemm
from triton_dist.profiler_utils import perf_func_with_l2_reset


def config_space():
    for bn, cg, gm, tma in itertools.product((256, 128, 64), (2, 1), (4, 8, 16), (True,)):
        yield GemmConfig(bn=bn, cta_group=cg, group_m=gm, use_tma_store=tma)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="+", default=["4096x4096x4096"]); ap.add_argument("--topk", type=int, default=3)
    ap.add_argument("--dtype", default="bfloat16"); ap.add_argument("--out", default="")
    args = ap.parse_args()
    dt = getattr(torch, args.dtype)
    results = {}
    for sh in args.shapes:
        M, N, K = (int(v) for v in sh.lower().split("x"))
        a, b = torch.randn(M, K, device="cuda", dtype=dt), torch.randn(N, K, device="cuda", dtype=dt)
        c = torch.empty(M, N, device="cuda", dtype=dt)
        rows = []
        for cfg in config_space():
            try:
                _, ms = perf_func_with_l2_reset(lambda: gemm(a, b, out=c, config=cfg), 10, 3)
                rows.append(dict(cfg=cfg.key(), ms=ms, tflops=2.0 * M * N * K / ms / 1e9))
            except Exception as e:      # noqa: BLE001
                rows.append(dict(cfg=cfg.key(), error=str(e)[:80]))
        rows.sort(key=lambda r: r.get("ms", 1e9))
        results[sh] = rows[:args.topk]
        print(sh, json.dumps(rows[:args.topk]))
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
