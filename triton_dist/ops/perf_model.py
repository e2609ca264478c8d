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
