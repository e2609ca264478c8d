"""Timing + profiling helpers (reference: /root/reference/python/triton_dist/profiler_utils.py:70-400)."""
from __future__ import annotations

import contextlib
import gzip
import json
import os
import time
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


def perf_func(func: Callable, iters: int, warmup_iters: int):
    """(last output, ms per call): CUDA events around ``iters`` back-to-back calls (profiler_utils.py:355-369)."""
    out = None
    for _ in range(warmup_iters):
        out = func()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = func()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) / iters
    t0 = time.perf_counter()
    for _ in range(iters):
        out = func()
    return out, (time.perf_counter() - t0) * 1e3 / iters


_L2_FLUSH = {}


def perf_func_with_l2_reset(func: Callable, iters: int, warmup_iters: int):
    """Like :func:`perf_func` but writes a 256 MB buffer (> the 126 MB L2 of a B200) before every timed call; only the
    calls themselves are inside the event pairs."""
    if not torch.cuda.is_available():
        return perf_func(func, iters, warmup_iters)
    buf = _L2_FLUSH.setdefault("b", torch.empty(256 << 20, dtype=torch.uint8, device="cuda"))
    out = None
    for _ in range(warmup_iters):
        out = func()
    total = 0.0
    for _ in range(iters):
        buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = func()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return out, total / iters


def max_over_ranks(ms: float, group=None) -> float:
    """Multi-GPU numbers are the slowest rank's device time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device="cuda" if torch.cuda.is_available() and dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def _merge_traces(traces: List[dict]) -> dict:
    events = []
    for rank, tr in enumerate(traces):
        for ev in tr.get("traceEvents", []):
            ev = dict(ev)
            if "pid" in ev:
                ev["pid"] = f"rank{rank}:{ev['pid']}"
            events.append(ev)
    return {"traceEvents": events, "displayTimeUnit": "ms"}


@contextlib.contextmanager
def group_profile(name: str = "trace", do_prof: bool = True, group=None, out_dir: str = "prof"):
    """torch.profiler on every rank -> chrome traces gathered to rank 0 and merged into one ``.json.gz``
    (profiler_utils.py:205-289)."""
    if not do_prof:
        yield None
        return
    acts = [torch.profiler.ProfilerActivity.CPU]
    if torch.cuda.is_available():
        acts.append(torch.profiler.ProfilerActivity.CUDA)
    with torch.profiler.profile(activities=acts, record_shapes=False) as prof:
        yield prof
    os.makedirs(out_dir, exist_ok=True)
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    tmp = os.path.join(out_dir, f"{name}_rank{rank}.json")
    prof.export_chrome_trace(tmp)
    trace = json.load(open(tmp))
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(trace, gathered, dst=0, group=group)
        if rank == 0:
            trace = _merge_traces(gathered)
    if rank == 0:
        with gzip.open(os.path.join(out_dir, f"{name}_merged.json.gz"), "wt") as f:
            json.dump(trace, f)


def print_benchmark_comparison(rows: List[dict], baseline_key: str = "torch", title: str = ""):
    """rows: [{"name":..., "<impl>": ms, ...}] -> aligned table with speedups over ``baseline_key``."""
    if not rows:
        return
    impls = [k for k in rows[0] if k != "name"]
    print(title)
    print(f"{'case':32s}" + "".join(f"{i:>14s}" for i in impls) + "".join(f"{'x' + i:>10s}" for i in impls if i != baseline_key))
    for r in rows:
        line = f"{r['name']:32s}" + "".join(f"{r[i]:14.4f}" for i in impls)
        line += "".join(f"{r[baseline_key] / r[i]:10.3f}" for i in impls if i != baseline_key)
        print(line)
