"""Rank the configurations recorded by ``tune_gemm --out results.json`` (reference: python/triton_dist/tools/tune/find_topk.py).
    python -m triton_dist.tools.tune.find_topk results.json --topk 3"""
import argparse
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("results"); ap.add_argument("--topk", type=int, default=3)
    args = ap.parse_args()
    data = json.load(open(args.results))
    wins = {}
    for shape, rows in data.items():
        rows = sorted((r for r in rows if "ms" in r), key=lambda r: r["ms"])[: args.topk]
        print(shape)
        for r in rows:
            print(f"  cfg={tuple(r['cfg'])}  {r['ms'] * 1e3:8.1f} us  {r['tflops']:7.1f} TFLOP/s")
            wins[tuple(r["cfg"])] = wins.get(tuple(r["cfg"]), 0) + 1
    print("configs by number of top-k appearances:")
    for cfg, n in sorted(wins.items(), key=lambda kv: -kv[1]):
        print(f"  {cfg}: {n}")


if __name__ == "__main__":
    main()
