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


def _intranode_gbps(world, *legacy) -> float:
    """The reference passes (local_world_size, intranode_bw_in_gbps, internode_bw_in_gbps) (comm_perf_model.py:94-131); here the second
    positional argument may be the bandwidth directly.  Everything is one NVSwitch domain, so only the intra-node number matters."""
    if len(legacy) >= 2:
        return float(legacy[1])
    if len(legacy) == 1:
        return float(legacy[0])
    return NVLINK_MEASURED_GBS


def estimate_all_gather_time_ms(nbytes_total: int, world: int, *legacy) -> float:
    return nbytes_total * (world - 1) / world / (_intranode_gbps(world, *legacy) * 1e9) * 1e3


def estimate_reduce_scatter_time_ms(nbytes_total: int, world: int, *legacy) -> float:
    return nbytes_total * (world - 1) / world / (_intranode_gbps(world, *legacy) * 1e9) * 1e3


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


# ------------------------------------------------------------------------------------------------------------
# Device table / tensor-core peak helpers in the reference's spelling (gemm_perf_model.py:49-235)
# ------------------------------------------------------------------------------------------------------------
_DEVICE_TABLE = {
    # name fragment: (dense bf16 TFLOP/s, dense fp8 TFLOP/s, HBM GB/s, SMs, boost MHz)
    "B200": (2250.0, 4500.0, 7700.0, 148, 1965),
    "B300": (2250.0, 4500.0, 7700.0, 148, 1965),
    "GB200": (2500.0, 5000.0, 8000.0, 148, 1965),
    "H100": (989.0, 1979.0, 3350.0, 132, 1830),
    "H800": (989.0, 1979.0, 3350.0, 132, 1830),
    "H200": (989.0, 1979.0, 4800.0, 132, 1830),
}


def _device_entry(device_name=None):
    if device_name is None:
        device_name = measured_peaks().get("gpu_name") or (torch.cuda.get_device_name() if torch.cuda.is_available() else "B200")
    for k in sorted(_DEVICE_TABLE, key=len, reverse=True):
        if k in str(device_name):
            return _DEVICE_TABLE[k]
    return _DEVICE_TABLE["B200"]


def is_fp8_dtype(dtype: torch.dtype) -> bool:
    return dtype in (torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e4m3fnuz, torch.float8_e5m2fnuz)


def get_tensorcore_tflops_by_device_name(dtype: torch.dtype, device_name=None) -> float:
    bf, f8, *_ = _device_entry(device_name)
    if is_fp8_dtype(dtype) or dtype == torch.int8:
        return f8
    if dtype == torch.float32:
        return bf / 2            # tf32
    return bf


def get_max_tensorcore_tflops(dtype: torch.dtype, clock_rate_mhz=None, device=None) -> float:
    """Nominal dense peak scaled to a clock (the reference derives it from SM count x per-SM rate x clock)."""
    entry = _device_entry(device)
    peak = get_tensorcore_tflops_by_device_name(dtype, device)
    return peak if not clock_rate_mhz else peak * float(clock_rate_mhz) / entry[4]


def get_full_tflops_approx(dtype: torch.dtype, device=None) -> float:
    return get_max_tensorcore_tflops(dtype, None, device)


def get_tflops_approx(device, num_ctas: int, num_warps: int, dtype: torch.dtype) -> float:
    """Share of the tensor-core peak a grid of ``num_ctas`` CTAs can reach (one tcgen05 issuer per CTA feeds one SM)."""
    sms = _device_entry(device)[3]
    return get_full_tflops_approx(dtype, device) * min(num_ctas, sms) / sms


def get_dram_gbps_by_device_name(device_name=None) -> float:
    return _device_entry(device_name)[2]


def get_device_multi_processor_count(device=None) -> int:
    return _device_entry(device)[3]


# ------------------------------------------------------------------------------------------------------------
# Tile-level GEMM model of csrc/gemm_sm100.cuh (waves of tiles on CTAs / CTA pairs)
# ------------------------------------------------------------------------------------------------------------
# sustained mainloop rate of ONE CTA in TFLOP/s by (cta_group, BN): fitted to profiles/README.md section 1 (cta_group 1) and the round-2
# plain-GEMM numbers (cta_group 2).  128-wide tiles re-read B twice as often per FLOP and are shared-memory bound.
_CTA_RATE = {(2, 256): TILE_RATE_TFLOPS_PER_PAIR / 2, (2, 192): TILE_RATE_TFLOPS_PER_PAIR / 2 * 0.97, (2, 128): 8.6,
             (1, 256): 10.0, (1, 192): 9.6, (1, 128): 6.3}
GEMM_PROLOGUE_US = 5.0                    # launch, tensormap prefetch, barrier init, TMEM alloc, first TMA round trip


def estimate_gemm_ms(M: int, N: int, K: int, cta_group: int = 2, bn: int = 256, sms: int = 148, split_k_tail: bool = True,
                     dtype: torch.dtype = torch.bfloat16) -> float:
    """What the persistent tcgen05 GEMM of this repo takes: ceil-waves of (128 * cta_group) x bn tiles over the CTAs (pairs), the last
    partial wave shortened by the split-K tail schedule; never below the HBM time of the operands."""
    tm = 128 * cta_group
    tiles = -(-M // tm) * -(-N // bn)
    workers = max(1, sms // cta_group)
    rate = _CTA_RATE.get((cta_group, bn), _CTA_RATE[(cta_group, 256)] * bn / 256) * cta_group
    if is_fp8_dtype(dtype):
        rate *= 1.6                       # measured MXFP8 / bf16 ratio of the same kernel (2.5 vs 1.55 PFLOP/s)
    tile_us = 2.0 * tm * bn * K / (rate * 1e12) * 1e6
    full, rem = divmod(tiles, workers)
    last = 0.0
    if rem:
        parts = min(4, workers // rem) if split_k_tail else 1
        last = 1.0 / max(parts, 1) + (0.08 if parts > 1 else 0.0)      # reduction of the partial accumulators
    es = torch.empty(0, dtype=dtype).element_size()
    t_mem = (M * K + N * K) * es / (get_dram_gbps() * 1e9) * 1e6
    return (max((full + last) * tile_us, t_mem) + GEMM_PROLOGUE_US) * 1e-3


def pick_gemm_config(M: int, N: int, K: int, sms: int = 148):
    """Model-cheapest (cta_group, bn) -- what ``tools/tune/tune_gemm.py`` finds by measurement."""
    best = None
    for cg, bn in ((2, 256), (2, 128), (1, 256), (1, 128)):
        t = estimate_gemm_ms(M, N, K, cg, bn, sms)
        if best is None or t < best[-1]:
            best = (cg, bn, t)
    return best


# ------------------------------------------------------------------------------------------------------------
# Collective models: t = alpha + beta * bytes, fitted at 8 x B200 (profiles/README.md section 3), scaled with the world size by the
# traffic each algorithm moves per rank
# ------------------------------------------------------------------------------------------------------------
# method: (alpha_us at W = 8, beta in us per byte at W = 8, traffic factor as a function of W used to rescale beta)
_AR_FIT = {
    "OneShot": (33.4, 1.40e-5, lambda w: (w - 1)),                  # every rank reads (W-1) full buffers over P2P
    "TwoShot": (26.5, 4.578e-6, lambda w: 2.0 * (w - 1) / w),       # reduce-scatter + all-gather over P2P
    "OneShot_Multimem": (21.6, 1.2225e-5, lambda w: float(w)),      # every rank ld_reduces the whole buffer: the switch reads W copies per rank
    "TwoShot_Multimem": (22.9, 3.704e-6, lambda w: 1.0),            # ld_reduce of 1/W + multimem.st broadcast: ~W-independent per rank
}


def estimate_allreduce_us(nbytes: int, world: int = 8, method: str = "TwoShot_Multimem") -> float:
    a, b, f = _AR_FIT[method]
    if world <= 1:
        return 5.0
    return a * (0.8 + 0.2 * world / 8.0) + b * f(world) / f(8) * nbytes


def pick_allreduce_method(nbytes: int, world: int = 8, multimem_ok: bool = True):
    names = [n for n in _AR_FIT if multimem_ok or "Multimem" not in n]
    best = min(names, key=lambda n: estimate_allreduce_us(nbytes, world, n))
    return best, estimate_allreduce_us(nbytes, world, best)


BARRIER_US = 9.0                          # cross-GPU flag barrier inside a kernel (release fence + one NVLink round trip)
COMM_KERNEL_LAUNCH_US = 6.0


def estimate_fast_allgather_us(shard_bytes: int, world: int = 8, mode: str = "push") -> float:
    """Small / medium message all-gather kernels (ops/comm.py fast_allgather): LL modes skip the barrier but move 2x the bytes."""
    if world <= 1:
        return COMM_KERNEL_LAUNCH_US
    ll = "ll" in mode
    mc = "multimem" in mode
    wire = shard_bytes * (2 if ll else 1)
    if mc:
        t = wire * world / (NVLS_MULTICAST_INGRESS_GBS * 1e9) * 1e6          # ingress of W shards; egress is 1x
    else:
        t = wire * (world - 1) / (NVLINK_ALLGATHER_PATTERN_GBS * 1e9) * 1e6
    return COMM_KERNEL_LAUNCH_US + t + (2.5 if ll else BARRIER_US)


def estimate_ep_dispatch_us(tokens_per_rank: int, hidden: int, topk: int, world: int = 8, bytes_per_elem: int = 1) -> float:
    """EP low-latency dispatch (csrc/ep_kernels.cu): every (token, k) row crosses NVLink once; fitted to 52.6 us (fp8) / 71.9 us (bf16)
    at 128 tokens x top-8 x 7168 on 8 ranks."""
    payload = tokens_per_rank * topk * hidden * bytes_per_elem * (world - 1) / max(world, 1)
    return 38.0 + payload / (NVLINK_ALLGATHER_PATTERN_GBS * 1e9) * 1e6 * (8.0 / 7.0)


def estimate_ep_combine_us(tokens_per_rank: int, hidden: int, topk: int, world: int = 8) -> float:
    """Combine returns bf16 rows and reduces top-k on the owner: 57.6 us (fp8 run: 110.2 - 52.6) / 70.5 us (bf16 run) measured."""
    payload = tokens_per_rank * topk * hidden * 2 * (world - 1) / max(world, 1)
    return 36.0 + payload / (NVLINK_ALLGATHER_PATTERN_GBS * 1e9) * 1e6 * (8.0 / 7.0)


# ---- peaks computed from the machine instead of looked up (reference: gemm_perf_model.py ``get_tensorcore_tflops_by_calc`` / ``get_simd_tflops``) --
# dense FLOP per clock per SM of the 5th-generation tensor core (tcgen05) by operand type, and of the CUDA cores
_TC_FLOP_PER_CLK_PER_SM = {"tf32": 4096, "bf16": 8192, "fp16": 8192, "fp8": 16384, "int8": 16384, "fp4": 32768}
_SIMD_FLOP_PER_CLK_PER_SM = {"fp64": 128, "fp32": 256, "fp16": 512, "bf16": 512}


def _dtype_key(dtype) -> str:
    if is_fp8_dtype(dtype):
        return "fp8"
    return {torch.float32: "tf32", torch.bfloat16: "bf16", torch.float16: "fp16", torch.int8: "int8", torch.uint8: "fp4"}.get(dtype, "bf16")


def get_tensorcore_dtype_support(device=None):
    """Operand types the tensor cores of sm_100a take (``kind::tf32 / f16 / f8f6f4 / i8 / mxf8f6f4 / mxf4``)."""
    return sorted(_TC_FLOP_PER_CLK_PER_SM)


def get_tensorcore_tflops_by_calc(dtype: torch.dtype, device=None, clock_rate_mhz=None) -> float:
    """SMs x FLOP / clk / SM x clock.  Without a GPU: 148 SMs at the 1.86 GHz the nominal 2.25 PFLOP/s (bf16) corresponds to."""
    sms = get_device_multi_processor_count(device) or 148
    mhz = clock_rate_mhz
    if mhz is None:
        from ..utils.topology import get_max_gpu_clock_rate_in_khz
        mhz = (get_max_gpu_clock_rate_in_khz(0) / 1e3) if torch.cuda.is_available() else 0.0
    mhz = mhz or 1860.0
    return sms * _TC_FLOP_PER_CLK_PER_SM[_dtype_key(dtype)] * mhz * 1e6 / 1e12


def get_simd_tflops(dtype: torch.dtype = torch.float32, device=None, clock_rate_mhz=None) -> float:
    """CUDA-core (FMA pipe) peak: what element-wise epilogues, softmax and the GEMV decode paths are bounded by when not by memory."""
    key = {torch.float64: "fp64", torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}.get(dtype, "fp32")
    sms = get_device_multi_processor_count(device) or 148
    return sms * _SIMD_FLOP_PER_CLK_PER_SM[key] * (clock_rate_mhz or 1860.0) * 1e6 / 1e12

