"""Topology / hardware queries (reference: /root/reference/python/triton_dist/nv_utils.py:88-318 via NVML).

A single HGX B200 box is one NVSwitch domain: every GPU reaches every peer over 18 NVLink-5 links
(900 GB/s per direction nominal), so the answers are mostly constants; NVML is consulted when importable."""
from __future__ import annotations

import os
from functools import lru_cache

import torch

from .. import _C

NVLINK5_GBPS_PER_DIR = 900.0
MEASURED_PEER_COPY_GBPS = 770.0      # /opt/skills/guides/B200_PROFILING.md


@lru_cache(None)
def get_device_info(dev: int = 0) -> dict:
    if not torch.cuda.is_available():
        return dict(sms=0, cc=(0, 0), multicast=False, smem_optin=0, l2_bytes=0)
    import ctypes as C
    out = (C.c_int * 8)()
    _C.check(_C.cuda_lib().td_device_info(dev, out), "td_device_info")
    return dict(sms=out[0], cc=(out[1], out[2]), multicast=bool(out[3]), smem_optin=out[4], l2_bytes=out[5],
                clock_khz=out[6], mem_clock_khz=out[7])


def _nvml():
    try:
        import pynvml
        pynvml.nvmlInit()
        return pynvml
    except Exception:
        return None


@lru_cache(None)
def has_fullmesh_nvlink() -> bool:
    if not torch.cuda.is_available():
        return False
    n = torch.cuda.device_count()
    if n <= 1:
        return True
    lib = _C.cuda_lib()
    return all(lib.td_can_access_peer(i, j) for i in range(n) for j in range(n) if i != j)


@lru_cache(None)
def get_nvlink_max_speed_gbps() -> float:
    nv = _nvml()
    if nv is None or not torch.cuda.is_available():
        return NVLINK5_GBPS_PER_DIR
    try:
        h = nv.nvmlDeviceGetHandleByIndex(0)
        total = 0.0
        for link in range(18):
            try:
                if nv.nvmlDeviceGetNvLinkState(h, link):
                    total += 50.0       # NVLink 5: 50 GB/s per direction per link
            except Exception:
                break
        return total or NVLINK5_GBPS_PER_DIR
    except Exception:
        return NVLINK5_GBPS_PER_DIR


def get_intranode_max_speed_gbps() -> float:
    return get_nvlink_max_speed_gbps() if has_fullmesh_nvlink() else 64.0


def get_numa_node(dev: int = 0) -> int:
    nv = _nvml()
    if nv is None:
        return 0
    try:
        h = nv.nvmlDeviceGetHandleByIndex(dev)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        p = f"/sys/bus/pci/devices/{bus.lower()[4:] if len(bus) > 12 else bus.lower()}/numa_node"
        return max(0, int(open(p).read())) if os.path.exists(p) else 0
    except Exception:
        return 0
