"""Choose the k tile configurations to ship: given the timings ``tune_gemm --out results.json`` recorded for many shapes, pick the k
configurations whose best-of-k is closest to the per-shape optimum over all shapes.

Reference: python/triton_dist/tools/tune/find_topk.py (``find_best_topk`` / ``find_best_topk_fast``: greedy selection over a slowdown
matrix read from autotune logs, shape filters ``IntFilter`` / ``parse_range``).  Same question, same greedy answer, numpy only:

    python -m triton_dist.tools.tune.find_topk results.json --topk 3 [--objective minimax] [-M 4096 | --M-range 128-8192-128] ...

prints the chosen configurations and the mean / p90 / p99 / max slowdown a heuristic restricted to them would pay, followed by the plain
per-shape ranking.  ``results.json``: ``{"(M, N, K)": [{"cfg": [...], "ms": float, "tflops": float}, ...], ...}``.
"""
from __future__ import annotations

import argparse
import ast
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


class IntFilter:
    """``None`` matches everything, an int matches itself, ``[lo, hi]`` matches the closed range."""

    def __init__(self, rule):
        if rule is not None and not isinstance(rule, int):
            if not (isinstance(rule, (list, tuple)) and len(rule) == 2 and all(isinstance(i, int) for i in rule) and rule[0] <= rule[1]):
                raise ValueError("rule must be None, an int or a [min, max] pair")
            rule = list(rule)
        self.rule = rule

    def match(self, val: int) -> bool:
        if self.rule is None:
            return True
        if isinstance(self.rule, int):
            return val == self.rule
        return self.rule[0] <= val <= self.rule[1]

    def is_int(self) -> bool:
        return isinstance(self.rule, int)

    def __repr__(self):
        return f"IntFilter(rule={self.rule})"


def parse_range(range_str: str) -> Tuple[int, int, int]:
    start, end, step = (int(x) for x in range_str.split("-"))
    assert start < end and step > 0
    return start, end, step


def parse_int_range_args(value: Optional[int], value_range: Optional[str]) -> IntFilter:
    if value:
        return IntFilter(int(value))
    if value_range:
        lo, hi, _ = parse_range(value_range)
        return IntFilter([lo, hi])
    return IntFilter(None)


def slowdown_matrix(data: Dict[str, List[dict]], filters: Sequence[IntFilter] = (IntFilter(None),) * 3):
    """-> (configs, shapes, S) with ``S[c, s] = time of config c on shape s / best time on shape s`` (inf where not measured)."""
    shapes, cfgs = [], []
    for key, rows in data.items():
        try:
            dims = tuple(int(x) for x in ast.literal_eval(key))[:3] if isinstance(key, str) else tuple(key)[:3]
        except (ValueError, SyntaxError):
            dims = ()
        if len(dims) == 3 and not all(f.match(d) for f, d in zip(filters, dims)):
            continue
        if any("ms" in r for r in rows):
            shapes.append(key)
            for r in rows:
                c = tuple(r["cfg"])
                if "ms" in r and c not in cfgs:
                    cfgs.append(c)
    S = np.full((len(cfgs), len(shapes)), np.inf)
    for j, key in enumerate(shapes):
        for r in data[key]:
            if "ms" in r:
                S[cfgs.index(tuple(r["cfg"])), j] = r["ms"]
        S[:, j] /= S[:, j].min()
    return cfgs, shapes, S


def find_best_topk(S: np.ndarray, topk: int = 1, objective: str = "mean", threshold: Optional[float] = 1.15):
    """The k rows minimising the score of the column-wise minima over the chosen rows (``mean``: average slowdown over shapes;
    ``minimax``: the worst shape): exhaustive when there are at most 50 000 subsets, greedy (as the reference) beyond that.  ``threshold`` drops configurations that are never within that
    factor of the optimum anywhere.  Returns (chosen row indices, (mean, p90, p99, max) of the final per-shape slowdowns)."""
    assert S.ndim == 2 and S.shape[0] > 0 and S.shape[1] > 0
    rows = [i for i in range(S.shape[0]) if threshold is None or S[i].min() <= threshold] or list(range(S.shape[0]))
    score = (lambda v: float(np.mean(v))) if objective == "mean" else (lambda v: float(np.max(v)))
    finite = lambda v: np.where(np.isfinite(v), v, 1e6)               # an unmeasured pair is a very bad pair, not a crash
    k = min(topk, len(rows))
    import itertools
    import math
    if math.comb(len(rows), k) <= 50000:
        # small enough to be exact (greedy picks the best generalist first and can miss a pair of complementary specialists)
        chosen = list(min(itertools.combinations(rows, k), key=lambda c: score(finite(S[list(c)].min(axis=0)))))
        best = S[chosen].min(axis=0)
    else:
        chosen: List[int] = []
        best = np.full(S.shape[1], np.inf)
        for _ in range(k):
            cand = min((r for r in rows if r not in chosen), key=lambda r: score(finite(np.minimum(best, S[r]))))
            chosen.append(cand)
            best = np.minimum(best, S[cand])
    b = np.where(np.isfinite(best), best, 1e6)
    return chosen, (float(b.mean()), float(np.quantile(b, 0.9)), float(np.quantile(b, 0.99)), float(b.max()))


find_best_topk_fast = find_best_topk


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("results")
    ap.add_argument("--topk", type=int, default=3)
    ap.add_argument("--objective", default="mean", choices=["mean", "minimax"])
    ap.add_argument("--threshold", type=float, default=1.15)
    for d in "MNK":
        ap.add_argument(f"-{d}", f"--{d}", type=int)
        ap.add_argument(f"--{d}-range", f"--{d}_range", type=str)
    a = ap.parse_args(argv)
    data = json.load(open(a.results))
    filters = [parse_int_range_args(getattr(a, d), getattr(a, f"{d}_range")) for d in "MNK"]
    cfgs, shapes, S = slowdown_matrix(data, filters)
    if not cfgs:
        print("no measured shapes match the filters")
        return 1
    chosen, (mean, p90, p99, worst) = find_best_topk(S, a.topk, a.objective, a.threshold)
    print(f"{len(shapes)} shapes, {len(cfgs)} configurations; best {len(chosen)} by {a.objective}:")
    for c in chosen:
        wins = int((S[c] == S[chosen].min(axis=0)).sum())
        print(f"  cfg={cfgs[c]}  best of the set on {wins} shapes, own mean slowdown {np.where(np.isfinite(S[c]), S[c], np.nan).mean():.3f}")
    print(f"slowdown vs per-shape optimum with this set: mean {mean:.3f}  p90 {p90:.3f}  p99 {p99:.3f}  max {worst:.3f}")
    for shape in shapes:
        rows = sorted((r for r in data[shape] if "ms" in r), key=lambda r: r["ms"])[: a.topk]
        print(shape)
        for r in rows:
            print(f"  cfg={tuple(r['cfg'])}  {r['ms'] * 1e3:8.1f} us" + (f"  {r['tflops']:7.1f} TFLOP/s" if "tflops" in r else ""))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
