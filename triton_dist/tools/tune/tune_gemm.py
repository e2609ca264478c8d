"""Offline GEMM config search (reference: python/triton_dist/tools/tune/tune_gemm.py, find_topk.py).
    python -m triton_dist.tools.tune.tune_gemm --shapes 4096x4096x4096 8192x8192x8192 --topk 3
    python -m triton_dist.tools.tune.tune_gemm --shapes 4096x12288x6144 --dry-run      # no GPU: rank by the tile-wave model
Every measured row also carries the tile-wave model's prediction (``ops/perf_model.estimate_gemm_ms``) so that a sweep doubles as a
calibration check of the model."""
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
    ap.add_argument("--dry-run", action="store_true", help="rank the configurations by the performance model only (no GPU needed)")
    args = ap.parse_args()
    dt = getattr(torch, args.dtype)
    from triton_dist.ops.perf_model import estimate_gemm_ms
    results = {}
    for sh in args.shapes:
        M, N, K = (int(v) for v in sh.lower().split("x"))
        if args.dry_run or not torch.cuda.is_available():
            rows = [dict(cfg=cfg.key(), model_ms=estimate_gemm_ms(M, N, K, cfg.cta_group, cfg.bn, dtype=dt)) for cfg in config_space()]
            rows.sort(key=lambda r: r["model_ms"])
            for r in rows:
                r["model_tflops"] = 2.0 * M * N * K / r["model_ms"] / 1e9
            for r in rows:
                r["ms"] = r["model_ms"]                      # a dry run's file ranks by the model (find_topk reads "ms")
            results[str((M, N, K))] = rows
            print(sh, "(model)", json.dumps(rows[:args.topk]))
            continue
        a, b = torch.randn(M, K, device="cuda", dtype=dt), torch.randn(N, K, device="cuda", dtype=dt)
        c = torch.empty(M, N, device="cuda", dtype=dt)
        rows = []
        for cfg in config_space():
            try:
                _, ms = perf_func_with_l2_reset(lambda: gemm(a, b, out=c, config=cfg), 10, 3)
                rows.append(dict(cfg=cfg.key(), ms=ms, tflops=2.0 * M * N * K / ms / 1e9,
                                 model_ms=estimate_gemm_ms(M, N, K, cfg.cta_group, cfg.bn, dtype=dt)))
            except Exception as e:      # noqa: BLE001
                rows.append(dict(cfg=cfg.key(), error=str(e)[:80]))
        rows.sort(key=lambda r: r.get("ms", 1e9))
        results[str((M, N, K))] = rows                  # every configuration: find_topk builds its slowdown matrix from the full table
        print(sh, json.dumps(rows[:args.topk]))
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
