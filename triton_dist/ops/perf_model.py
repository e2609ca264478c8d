"""Speed-of-light models for B200 (reference: kernels/nvidia/gemm_perf_model.py:143-235, comm_perf_model.py:116 -- device
table without a B200 entry).  Denominators come from MEASURED_PEAKS.json when present (driver-measured on this pool)."""
from __future__ import annotations

import json
import os
from functools import lru_cache

import torch

_NOMINAL = {"bf16_tflops": 2250.0, "fp8_tflops": 4500.0, "hbm_gbs": 7700.0, "nvlink_gbs": 900.0}
_FALLBACK = {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
NVLINK_MEASURED_GBS = 770.0          # peer copy, per direction (B200_PROFILING.md)
ALLREDUCE_BUSBW_MEASURED_GBS = 725.0


@lru_cache(None)
def measured_peaks() -> dict:
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        return json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    except Exception:
        return dict(_FALLBACK, source="fallback")


def get_tensorcore_tflops(dtype: torch.dtype, measured: bool = True) -> float:
    if measured and dtype in (torch.bfloat16, torch.float16):
        return float(measured_peaks().get("bf16_tflops", _FALLBACK["bf16_tflops"]))
    if dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
        return 2.0 * get_tensorcore_tflops(torch.bfloat16, measured) if measured else _NOMINAL["fp8_tflops"]
    return _NOMINAL["bf16_tflops"]


def get_dram_gbps(measured: bool = True) -> float:
    return float(measured_peaks().get("hbm_gbs", _FALLBACK["hbm_gbs"])) if measured else _NOMINAL["hbm_gbs"]


def estimate_gemm_sol_time_ms(M: int, N: int, K: int, dtype: torch.dtype = torch.bfloat16) -> float:
    es = torch.empty(0, dtype=dtype).element_size()
    t_flop = 2.0 * M * N * K / (get_tensorcore_tflops(dtype) * 1e12)
    t_mem = (M * K + N * K + M * N) * es / (get_dram_gbps() * 1e9)
    return max(t_flop, t_mem) * 1e3


def get_nic_gbps_per_gpu() -> float:
    return 0.0       # single NVSwitch domain


def estimate_all_gather_time_ms(nbytes_total: int, world: int, gbps: float = NVLINK_MEASURED_GBS) -> float:
    return nbytes_total * (world - 1) / world / (gbps * 1e9) * 1e3


def estimate_reduce_scatter_time_ms(nbytes_total: int, world: int, gbps: float = NVLINK_MEASURED_GBS) -> float:
    return nbytes_total * (world - 1) / world / (gbps * 1e9) * 1e3


def fused_roofline_ms(flops_per_rank: float, nvlink_bytes_per_rank: float, dtype=torch.bfloat16) -> float:
    """Target of a fused compute+collective kernel: the slower of compute at the measured GEMM peak and the bytes that
    must cross one NVLink port at the measured link bandwidth."""
    return max(flops_per_rank / (get_tensorcore_tflops(dtype) * 1e12), nvlink_bytes_per_rank / (NVLINK_MEASURED_GBS * 1e9)) * 1e3


# ------------------------------------------------------------------------------------------------------------
# Calibrated models of the fused kernels (constants measured on 8xB200 in round 2: profiles/README.md, round-2 section,
# intra-kernel traces profiles/r2/ag_gemm_trace_n8_*.json.gz).  They predict what the kernels in csrc/gemm_sm100.cuh actually
# do, not a speed of light: used to pick configurations without running a sweep and to sanity-check measurements.
# ------------------------------------------------------------------------------------------------------------
NVLINK_ALLGATHER_PATTERN_GBS = 510.0     # per direction per GPU when every GPU sends to and receives from all peers (SM stores)
NVLS_MULTICAST_INGRESS_GBS = 380.0       # multimem.st all-gather: ingress per GPU
RELEASE_FENCE_US = 7.0                   # fence.acq_rel.sys after NVLink stores: pure latency, independent of bytes in flight
KERNEL_LAUNCH_GAP_US = 12.0              # back-to-back fused launches: launch + prologue (TMEM alloc, barrier init, tensormaps)
SM_PUSH_GBS = 42.0                       # one SM, 256 threads of 16-byte stores over NVLink
TILE_RATE_TFLOPS_PER_PAIR = 1676.7 / 74  # one CTA pair of the 2-CTA 256x256 kernel at the measured cuBLAS-level rate


def mainloop_us(tm: int, bn: int, K: int, pairs_rate_tflops: float = TILE_RATE_TFLOPS_PER_PAIR) -> float:
    """Mainloop time of ONE tm x bn tile over K on one CTA pair (tm = 256) / one CTA (tm = 128: half the rate)."""
    rate = pairs_rate_tflops * (tm / 256.0) * (1.0 if tm == 256 else 0.5)      # a single CTA is shared-memory bound (2x slower per FLOP)
    return 2.0 * tm * bn * K / (rate * 1e12) * 1e6


def estimate_ag_gemm_ms(M: int, N_local: int, K: int, world: int, transport: str = "sm_k", kslices: int = 2, groups: int = 1,
                        n_comm: int = 32, sms: int = 148, esz: int = 2) -> float:
    """AllGather + GEMM with the K-sliced transports: transfer of (W-1) shards at the all-gather fabric rate, one release fence per
    round of slices, the MMAs of the last round after the last byte, epilogue, launch gap; never below the GEMM itself."""
    if world <= 1:
        return estimate_gemm_sol_time_ms(M, N_local, K)
    shard = (M // world) * K * esz
    if transport == "multicast":
        xfer = world * shard / (NVLS_MULTICAST_INGRESS_GBS * 1e9) * 1e6
    else:
        rate = min(NVLINK_ALLGATHER_PATTERN_GBS, n_comm * SM_PUSH_GBS)
        xfer = (world - 1) * shard / (rate * 1e9) * 1e6
    rounds = max(1, (kslices + groups - 1) // max(groups, 1))
    tile = mainloop_us(256, 256, K)
    tail = tile * min(1.0, groups / max(kslices, 1))
    gemm_tiles = -(-M // 256) * -(-N_local // 256)
    workers = max(1, (sms - n_comm) // 2)
    gemm_only = -(-gemm_tiles // workers) * tile
    t = max(xfer + RELEASE_FENCE_US * min(rounds, 2) + tail, gemm_only) + 3.0 + KERNEL_LAUNCH_GAP_US
    return t * 1e-3


def estimate_gemm_rs_ms(M: int, N: int, K_local: int, world: int, sms: int = 148, esz: int = 2, split_k_tail: bool = True) -> float:
    """GEMM + ring ReduceScatter fused in the epilogue: the ring traffic hides behind the mainloop, so the model is the wave count of
    256x256 tiles on the CTA pairs (the split-K tail turns the last partial wave into 1 / parts of a wave) + one ring hop."""
    tiles = -(-M // 256) * -(-N // 256)
    workers = sms // 2
    full, rem = divmod(tiles, workers)
    last = 0.0
    if rem:
        parts = min(4, workers // rem) if split_k_tail else 1
        last = 1.0 / max(parts, 1)
    tile = mainloop_us(256, 256, K_local)
    ring_bytes = (world - 1) / world * M * N * esz
    ring = ring_bytes / (NVLINK_MEASURED_GBS * 1e9) * 1e6
    hop = 256 * 256 * esz / (SM_PUSH_GBS * 1e9) * 1e6 * 2 + RELEASE_FENCE_US if world > 1 else 0.0
    return (max((full + last) * tile, ring) + hop + KERNEL_LAUNCH_GAP_US) * 1e-3


def pick_ag_transport(M: int, N_local: int, K: int, world: int, multicast_ok: bool = True):
    """Cheapest (model) K-sliced configuration for ``ag_gemm``: returns (transport, kslices, groups, n_comm, predicted ms)."""
    best = None
    for tr in (("sm_k", "multicast") if multicast_ok else ("sm_k",)):
        for ks, gr, nc in ((2, 1, 32), (4, 2, 32), (6, 2, 32), (8, 3, 24)):
            t = estimate_ag_gemm_ms(M, N_local, K, world, tr, ks, gr, nc)
            if best is None or t < best[-1]:
                best = (tr, ks, gr, nc, t)
    return best
