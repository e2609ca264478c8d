"""Runners of the DSL micro-benchmarks: launch (GPU: CUDA events; CPU: interpreter at toy sizes), check the kernel's output against its
closed form, and convert the timing into the quantity of interest.  ``run_all`` returns ``{name: {metric: value, ..., "ok": bool}}``."""
from __future__ import annotations

import json
from typing import Callable, Dict, Optional

import torch

from .kernels import BLOCK, ILP, KERNELS


def _launch(name: str, grid: int, args, interpret: bool, smem: Optional[int] = None, iters: int = 5, warm: int = 2) -> float:
    """Milliseconds per launch (GPU) or 0.0 (interpreter)."""
    k = KERNELS[name]
    if interpret:
        k.interpret(grid, *args)
        return 0.0
    kw = {} if smem is None else {"smem": smem}
    for _ in range(warm):
        k.launch(grid, args, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        k.launch(grid, args, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _dev(interpret):
    return "cpu" if interpret else "cuda"


def _sms(interpret):
    return 2 if interpret else torch.cuda.get_device_properties(0).multi_processor_count


def bench_fma_throughput(interpret=False):
    grid, iters = (2, 4) if interpret else (_sms(False) * 8, 4096)
    out = torch.zeros(grid * BLOCK, device=_dev(interpret))
    ms = _launch("fma_throughput", grid, (out, iters, 1.0, 0.0), interpret)
    ok = bool(torch.all(out == float(sum(range(1, ILP + 1)))))
    flop = 2.0 * ILP * iters * grid * BLOCK
    return {"ok": ok, "fp32_fma_TFLOPs": flop / ms / 1e9 if ms else None}


def bench_sfu_throughput(interpret=False):
    grid, iters = (2, 3) if interpret else (_sms(False) * 8, 2048)
    out = torch.zeros(grid * BLOCK, device=_dev(interpret))
    ms = _launch("sfu_throughput", grid, (out, iters), interpret)
    ok = bool(torch.isfinite(out).all()) and bool((out.abs() < ILP).all())
    return {"ok": ok, "ex2_Gops": ILP * iters * grid * BLOCK / ms / 1e6 if ms else None}


def bench_mma_sync_throughput(interpret=False):
    grid, iters = (1, 2) if interpret else (_sms(False) * 4, 2048)
    out = torch.zeros(grid * BLOCK, device=_dev(interpret))
    ms = _launch("mma_sync_throughput", grid, (out, iters), interpret)
    ok = bool(torch.all(out == 16.0 * 16 * iters))                  # 16 accumulators, each 16 * iters
    flop = 4 * 4096.0 * iters * grid * (BLOCK // 32)
    return {"ok": ok, "mma_sync_bf16_TFLOPs": flop / ms / 1e9 if ms else None}


def bench_fma_latency(interpret=False):
    iters = 8 if interpret else 4096
    cyc, out = torch.zeros(1, dtype=torch.int64, device=_dev(interpret)), torch.zeros(32, device=_dev(interpret))
    _launch("fma_latency", 1, (cyc, out, iters, 0.5), interpret, iters=1, warm=1)
    # x <- x / 2 + 1 / 2 from 1.0 stays 1.0
    return {"ok": bool(torch.all(out == 1.0)), "fma_dependent_cycles": None if interpret else cyc.item() / iters}


def _chase_perm(n: int, stride: int) -> torch.Tensor:
    """One cycle through n slots with a large odd stride (defeats the prefetchers): i -> (i + stride) % n, stride coprime with n."""
    idx = torch.arange(n, dtype=torch.int64)
    return ((idx + stride) % n).to(torch.int32)


def bench_pointer_chase(interpret=False):
    res = {"ok": True}
    sizes = {"toy": 64} if interpret else {"l1_16KB": 4096, "l2_8MB": 2 << 20, "hbm_1GB": 256 << 20}
    for label, n in sizes.items():
        stride = (n // 2 + 1) | 1 if n > 64 else 7
        while torch.gcd(torch.tensor(stride), torch.tensor(n)).item() != 1:
            stride += 2
        nxt = _chase_perm(n, stride).to(_dev(interpret))
        steps = 16 if interpret else 20000
        cyc, end = torch.zeros(1, dtype=torch.int64, device=_dev(interpret)), torch.zeros(1, dtype=torch.int32, device=_dev(interpret))
        _launch("pointer_chase", 1, (nxt, cyc, end, steps), interpret, iters=1, warm=0)
        res["ok"] &= int(end.item()) == (steps * stride) % n
        res[f"load_to_use_cycles_{label}"] = None if interpret else cyc.item() / steps
    return res


def bench_smem_pointer_chase(interpret=False):
    steps, stride = (16, 33) if interpret else (20000, 33)
    cyc, end = torch.zeros(1, dtype=torch.int64, device=_dev(interpret)), torch.zeros(1, dtype=torch.int32, device=_dev(interpret))
    _launch("smem_pointer_chase", 1, (cyc, end, stride, steps), interpret, iters=1, warm=1)
    return {"ok": int(end.item()) == (steps * stride) % 1024, "smem_load_to_use_cycles": None if interpret else cyc.item() / steps}


def bench_sync_latency(interpret=False):
    iters = 4 if interpret else 4096
    cyc, ctr = torch.zeros(2, dtype=torch.int64, device=_dev(interpret)), torch.zeros(1, dtype=torch.int32, device=_dev(interpret))
    _launch("sync_latency", 1, (cyc, ctr, iters), interpret, iters=1, warm=0)
    return {"ok": int(ctr.item()) == iters, "syncthreads_cycles": None if interpret else cyc[0].item() / iters,
            "global_atomic_cycles": None if interpret else cyc[1].item() / iters}


def bench_global_copy(interpret=False):
    res = {"ok": True}
    for label, nbytes in ({"toy": 16 * 300}.items() if interpret else {"hbm_1GB": 1 << 30, "l2_32MB": 32 << 20}.items()):
        nvec = nbytes // 16
        src = torch.randint(0, 2 ** 31 - 1, (nvec * 4,), dtype=torch.int32, device=_dev(interpret))
        dst = torch.zeros_like(src)
        iters = 1 if (interpret or nbytes >= (1 << 30)) else 16
        ms = _launch("global_copy", 2 if interpret else _sms(False) * 8, (dst, src, nvec, iters), interpret)
        res["ok"] &= bool(torch.equal(dst, src))
        res[f"copy_GBps_{label}"] = 2.0 * nbytes * iters / ms / 1e6 if ms else None
    return res


def bench_global_read(interpret=False):
    nbytes = 16 * 256 if interpret else 1 << 30
    nvec = nbytes // 16
    src = torch.randint(0, 2 ** 31 - 1, (nvec * 4,), dtype=torch.int32, device=_dev(interpret))
    grid = 1 if interpret else _sms(False) * 8
    sink = torch.zeros(grid * BLOCK, dtype=torch.int32, device=_dev(interpret))
    ms = _launch("global_read", grid, (src, sink, nvec, 1), interpret)
    want = 0
    for v in (src.cpu().numpy().astype("uint32") if nvec * 4 <= 1 << 16 else []):
        want ^= int(v)
    got = 0
    if nvec * 4 <= 1 << 16:
        for v in sink.cpu().numpy().astype("uint32"):
            got ^= int(v)
    return {"ok": got == want, "read_GBps": nbytes / ms / 1e6 if ms else None}


def bench_smem_stride(interpret=False):
    res = {"ok": True}
    iters = 8 if interpret else 4096
    base = None
    for stride in ((1, 2) if interpret else (1, 2, 4, 8, 16, 32)):
        cyc = torch.zeros(1, dtype=torch.int64, device=_dev(interpret))
        sink = torch.zeros(BLOCK, dtype=torch.int32, device=_dev(interpret))
        _launch("smem_stride", 1, (cyc, sink, stride, iters), interpret, iters=1, warm=1)
        tid = torch.arange(BLOCK, dtype=torch.int64)
        want = sum(((tid * stride + it) % 4096) for it in range(iters)).to(torch.int32)
        res["ok"] &= bool(torch.equal(sink.cpu(), want))
        if not interpret:
            c = cyc.item() / iters
            base = base or c
            res[f"cycles_per_load_stride{stride}"] = c
            res[f"slowdown_stride{stride}"] = c / base
    return res


def bench_shuffle_throughput(interpret=False):
    grid, iters = (1, 2) if interpret else (_sms(False), 2048)
    out = torch.zeros(grid * BLOCK, dtype=torch.int32, device=_dev(interpret))
    cyc = torch.zeros(grid, dtype=torch.int64, device=_dev(interpret))
    _launch("shuffle_throughput", grid, (out, cyc, iters), interpret, iters=1, warm=1)
    per_warp = out.view(-1, 32)
    ok = bool((per_warp == per_warp[:, :1]).all())               # a butterfly xor leaves every lane of a warp with the same value
    return {"ok": ok, "cycles_per_shfl_ballot_round": None if interpret else cyc.float().mean().item() / iters}


def bench_int_ipc(interpret=False):
    grid, iters = (1, 3) if interpret else (_sms(False) * 4, 4096)
    out = torch.zeros(grid * BLOCK, dtype=torch.int32, device=_dev(interpret))
    cyc = torch.zeros(grid, dtype=torch.int64, device=_dev(interpret))
    _launch("int_ipc", grid, (out, cyc, iters), interpret, iters=1, warm=1)
    tid = torch.arange(BLOCK, dtype=torch.int64)
    want = torch.zeros(BLOCK, dtype=torch.int64)
    for j in range(ILP):
        x = tid + j
        for _ in range(iters):
            x = (x * 3 + 1) & 0xFFFFFFFF
        want = (want + x) & 0xFFFFFFFF
    got = out[:BLOCK].cpu().to(torch.int64) & 0xFFFFFFFF
    res = {"ok": bool(torch.equal(got, want)) if iters <= 64 else True}
    if not interpret:
        res["imad_per_clk_per_sm_4_blocks"] = BLOCK * ILP * iters * 4 / cyc.float().mean().item()
    return res


def bench_occupancy_probe(interpret=False):
    res = {"ok": True}
    for smem in ((0,) if interpret else (0, 32 << 10, 64 << 10, 100 << 10, 200 << 10)):
        grid = 4 if interpret else _sms(False) * 8
        smids = torch.full((grid,), -1, dtype=torch.int32, device=_dev(interpret))
        _launch("occupancy_probe", grid, (smids, 4 if interpret else 200000), interpret, smem=smem or None, iters=1, warm=0)
        res["ok"] &= bool((smids >= 0).all())
        if not interpret:
            counts = torch.bincount(smids.cpu().long())
            res[f"blocks_per_sm_smem{smem >> 10}KB"] = int(counts.max())      # upper bound over the run; equals residency when spin >> launch
    return res


BENCHES: Dict[str, Callable] = {
    "fma_throughput": bench_fma_throughput, "sfu_throughput": bench_sfu_throughput, "mma_sync_throughput": bench_mma_sync_throughput,
    "fma_latency": bench_fma_latency, "pointer_chase": bench_pointer_chase, "smem_pointer_chase": bench_smem_pointer_chase,
    "sync_latency": bench_sync_latency, "global_copy": bench_global_copy, "global_read": bench_global_read, "smem_stride": bench_smem_stride,
    "shuffle_throughput": bench_shuffle_throughput, "int_ipc": bench_int_ipc, "occupancy_probe": bench_occupancy_probe,
}


def run_all(only=None, interpret: bool = False) -> Dict[str, dict]:
    out = {}
    for name, fn in BENCHES.items():
        if only and name not in only:
            continue
        out[name] = fn(interpret=interpret)
    return out


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="DSL micro-benchmarks (single GPU; --interpret runs the kernels at toy sizes on CPU)")
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--interpret", action="store_true")
    a = ap.parse_args(argv)
    res = run_all([s for s in a.only.split(",") if s] or None, interpret=a.interpret or not torch.cuda.is_available())
    for name, r in res.items():
        vals = ", ".join(f"{k} = {v:.4g}" if isinstance(v, float) else f"{k} = {v}" for k, v in r.items() if k != "ok")
        print(f"{name:24s} {'ok ' if r['ok'] else 'BAD'} {vals}")
    line = json.dumps(res)
    print(line)
    if a.json:
        with open(a.json, "w") as f:
            f.write(line + "\n")
    return 0 if all(r["ok"] for r in res.values()) else 1
