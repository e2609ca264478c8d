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


# ------------------------------------------------------------------------------------------------------------
# trace utilities of the reference (profiler_utils.py: load_json / process_trace_json / ParallelJsonDumper / get_torch_prof_ctx /
# AutoExportProfiler / benchmark_latency_memory)
# ------------------------------------------------------------------------------------------------------------
def load_json(path: str) -> dict:
    """A chrome trace from ``.json`` or ``.json.gz``."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        return json.load(f)


def process_trace_json(trace: dict, rank: int, pid_stride: int = 1_000_000) -> dict:
    """Make the events of one rank mergeable with other ranks': numeric pids are shifted by ``rank * pid_stride`` (stream / thread
    lanes of different ranks must not collide), process-name metadata gets a ``rank N`` prefix."""
    out = []
    for ev in trace.get("traceEvents", []):
        ev = dict(ev)
        pid = ev.get("pid")
        if isinstance(pid, int):
            ev["pid"] = pid + rank * pid_stride
        elif pid is not None:
            ev["pid"] = f"rank{rank}:{pid}"
        if ev.get("ph") == "M" and ev.get("name") == "process_name" and isinstance(ev.get("args"), dict):
            ev["args"] = dict(ev["args"], name=f"rank {rank} " + str(ev["args"].get("name", "")))
        out.append(ev)
    return {"traceEvents": out, "displayTimeUnit": trace.get("displayTimeUnit", "ms")}


class ParallelJsonDumper:
    """Write a (large) merged trace as gzip'ed JSON with the event list serialised by a thread pool (json.dumps releases no GIL, but
    gzip does: chunks are compressed concurrently and concatenated -- concatenated gzip members form one valid gzip stream)."""

    def __init__(self, workers: int = 4, chunk_events: int = 50_000):
        self.workers, self.chunk_events = max(1, workers), chunk_events

    def dump(self, trace: dict, path: str):
        import concurrent.futures as cf
        evs = trace.get("traceEvents", [])
        chunks = [evs[i:i + self.chunk_events] for i in range(0, len(evs), self.chunk_events)] or [[]]

        def enc(i_chunk):
            i, chunk = i_chunk
            body = ",".join(json.dumps(e, separators=(",", ":")) for e in chunk)
            return gzip.compress((("," if i and body else "") + body).encode())
        head = gzip.compress(('{"displayTimeUnit":"%s","traceEvents":[' % trace.get("displayTimeUnit", "ms")).encode())
        with cf.ThreadPoolExecutor(self.workers) as ex:
            parts = list(ex.map(enc, enumerate(chunks)))
        with open(path, "wb") as f:
            f.write(head)
            for part in parts:
                f.write(part)
            f.write(gzip.compress(b"]}"))


def get_torch_prof_ctx(do_prof: bool):
    """``with get_torch_prof_ctx(flag) as prof`` -- a torch profiler (CPU + CUDA) or a null context."""
    if not do_prof:
        return contextlib.nullcontext()
    acts = [torch.profiler.ProfilerActivity.CPU] + ([torch.profiler.ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
    return torch.profiler.profile(activities=acts, record_shapes=True, with_stack=False)


class AutoExportProfiler:
    """Profile a region and export on exit: one chrome trace per rank, merged on rank 0 when a process group is up."""

    def __init__(self, name: str = "trace", out_dir: str = "prof", group=None, merge: bool = True):
        self.name, self.out_dir, self.group, self.merge = name, out_dir, group, merge
        self.merged_path = None

    def __enter__(self):
        self._ctx = group_profile(self.name, True, self.group, self.out_dir) if self.merge else get_torch_prof_ctx(True)
        self.prof = self._ctx.__enter__()
        return self.prof

    def __exit__(self, *exc):
        r = self._ctx.__exit__(*exc)
        rank = dist.get_rank() if dist.is_initialized() else 0
        if self.merge:
            self.merged_path = os.path.join(self.out_dir, f"{self.name}_merged.json.gz") if rank == 0 else None
        else:
            os.makedirs(self.out_dir, exist_ok=True)
            self.merged_path = os.path.join(self.out_dir, f"{self.name}_rank{rank}.json")
            self.prof.export_chrome_trace(self.merged_path)
        return r


def benchmark_latency_memory(func: Callable, iters: int = 10, warmup_iters: int = 3):
    """(result, ms per call, peak device memory in bytes during the timed calls)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
    out, ms = perf_func(func, iters, warmup_iters)
    peak = torch.cuda.max_memory_allocated() if torch.cuda.is_available() else 0
    return out, ms, peak


from .utils import sleep_async  # noqa: E402,F401  (the reference exports it from here too)
