"""Intra-kernel profiler: host side (reference: /root/reference/python/triton_dist/tools/profiler/{context,viewer}.py).

Device side: csrc/td/profiler.cuh (``td::prof_record(buf, slot, tag, is_start)`` -> one u64 per event).  Kernels take
the buffer through their argument struct; pass ``ProfilerBuffer.attach(args)`` before launching.  Export produces a
Chrome/Perfetto ``traceEvents`` JSON with one process per rank and one thread per (CTA, warp) slot.
"""
from __future__ import annotations

import gzip
import json
import os
from typing import Dict, List, Optional

import torch

GEMM_TASK_NAMES = {1: "ag_push(issue)", 2: "ag_push(wait complete)", 3: "producer flag wait", 4: "mainloop tile", 5: "epilogue tile", 6: "ag flag publish"}


class ProfilerBuffer:
    def __init__(self, max_num_profile_slots: int = 148 * 8, cap: int = 256, trace_file: Optional[str] = None,
                 task_names: Optional[Dict[int, str]] = None, device=None):
        self.num_slots, self.cap = max_num_profile_slots, cap
        self.trace_file = trace_file
        self.task_names = dict(task_names or GEMM_TASK_NAMES)
        self.buf = torch.zeros((self.num_slots, cap), dtype=torch.int64, device=device or ("cuda" if torch.cuda.is_available() else "cpu"))

    def reset(self):
        self.buf.zero_()

    def attach(self, args):
        """Fill the ``prof_*`` fields of a ctypes launch-argument struct."""
        args.prof_buf, args.prof_cap, args.prof_slots = self.buf.data_ptr(), self.cap, self.num_slots

    def events(self) -> List[dict]:
        b = self.buf.cpu()
        out = []
        for slot in range(self.num_slots):
            n = int(b[slot, 0])
            for i in range(1, n + 1):
                v = int(b[slot, i]) & 0xFFFFFFFFFFFFFFFF
                out.append(dict(slot=slot, tag=(v >> 56) & 0xFF, start=bool((v >> 55) & 1), ns=v & ((1 << 55) - 1)))
        return out

    def __enter__(self):
        self.reset()
        return self

    def __exit__(self, *exc):
        if self.trace_file:
            export_to_perfetto_trace(self, self.trace_file)


def alloc_profiler_buffer(max_num_profile_slots: int = 148 * 8, cap: int = 256, **kw) -> ProfilerBuffer:
    return ProfilerBuffer(max_num_profile_slots, cap, **kw)


def reset_profiler_buffer(pb: ProfilerBuffer):
    pb.reset()


def export_to_perfetto_trace(pb: ProfilerBuffer, path: str, rank: int = 0, warps_per_cta: int = 8) -> str:
    """Write a Chrome-trace JSON (load in ui.perfetto.dev / chrome://tracing): B/E events per (CTA, warp)."""
    evs = pb.events()
    t0 = min((e["ns"] for e in evs), default=0)
    trace = []
    for e in evs:
        trace.append({"name": pb.task_names.get(e["tag"], f"task{e['tag']}"), "ph": "B" if e["start"] else "E",
                      "ts": (e["ns"] - t0) / 1e3, "pid": f"rank{rank}", "tid": f"cta{e['slot'] // warps_per_cta}.w{e['slot'] % warps_per_cta}"})
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as f:
        json.dump({"traceEvents": trace, "displayTimeUnit": "ns"}, f)
    return path


def summarize(pb: ProfilerBuffer) -> Dict[str, dict]:
    """Per-tag duration statistics in microseconds (pairs each start with the next end of the same slot/tag)."""
    open_ev, durs = {}, {}
    for e in pb.events():
        k = (e["slot"], e["tag"])
        if e["start"]:
            open_ev[k] = e["ns"]
        elif k in open_ev:
            durs.setdefault(e["tag"], []).append((e["ns"] - open_ev.pop(k)) / 1e3)
    out = {}
    for tag, d in durs.items():
        d.sort()
        out[pb.task_names.get(tag, str(tag))] = dict(n=len(d), mean_us=sum(d) / len(d), p50_us=d[len(d) // 2], max_us=d[-1], total_us=sum(d))
    return out


Profiler = ProfilerBuffer


def decode_tag(v: int) -> dict:
    """One 64-bit event word of ``td/profiler.cuh`` -> ``{tag, start, ns}`` (tag: 8 bits, begin / end: 1 bit, globaltimer: 55 bits)."""
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    return dict(tag=(v >> 56) & 0xFF, start=bool((v >> 55) & 1), ns=v & ((1 << 55) - 1))


def parse_to_tracks(pb: ProfilerBuffer, warps_per_cta: int = 8) -> Dict[str, List[dict]]:
    """Events grouped per (CTA, warp) track and paired into intervals: ``{"cta3.w1": [{"name", "tag", "start_us", "dur_us"}, ...]}``
    (reference: tools/profiler/viewer.py ``parse_to_tracks``).  Unmatched begins (a kernel that trapped) are kept with ``dur_us = None``."""
    evs = pb.events()
    t0 = min((e["ns"] for e in evs), default=0)
    tracks: Dict[str, List[dict]] = {}
    open_: Dict[tuple, List[dict]] = {}
    for e in sorted(evs, key=lambda e: (e["slot"], e["ns"])):
        track = f"cta{e['slot'] // warps_per_cta}.w{e['slot'] % warps_per_cta}"
        key = (e["slot"], e["tag"])
        if e["start"]:
            rec = dict(name=pb.task_names.get(e["tag"], f"task{e['tag']}"), tag=e["tag"], start_us=(e["ns"] - t0) / 1e3, dur_us=None)
            tracks.setdefault(track, []).append(rec)
            open_.setdefault(key, []).append(rec)
        elif open_.get(key):
            rec = open_[key].pop()
            rec["dur_us"] = (e["ns"] - t0) / 1e3 - rec["start_us"]
    return tracks


_EXPORT_TRACE = [False]


def set_export_trace_on():
    """Ops that accept ``profiler=`` export their trace on completion while this is on (reference: tools/profiler/context.py)."""
    _EXPORT_TRACE[0] = True


def set_export_trace_off():
    _EXPORT_TRACE[0] = False


def get_export_trace_on() -> bool:
    return _EXPORT_TRACE[0]

