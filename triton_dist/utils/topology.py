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


# ---- NVML helpers under the reference's names (nv_utils.py:52-431); every one degrades to a neutral answer without NVML / a GPU ----------
def ensure_nvml_initialized() -> bool:
    return _nvml() is not None


def with_pynvml() -> bool:
    try:
        import pynvml  # noqa: F401
        return True
    except Exception:
        return False


def _handle(device_id=None):
    nv = _nvml()
    if nv is None:
        return None, None
    try:
        idx = torch.cuda.current_device() if device_id is None and torch.cuda.is_available() else int(device_id or 0)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(t.strip().isdigit() for t in vis.split(",")) and idx < len(vis.split(",")):
            idx = int(vis.split(",")[idx])                   # NVML numbers physical devices
        return nv, nv.nvmlDeviceGetHandleByIndex(idx)
    except Exception:
        return nv, None


def nvsmi(attrs, device_id: int = 0, dtype: type = int):
    """``nvidia-smi --query-gpu=<attrs>`` for one device -> list of parsed values (empty when nvidia-smi is unavailable)."""
    import shutil
    import subprocess
    exe = shutil.which("nvidia-smi")
    if exe is None:
        return []
    attrs = [attrs] if isinstance(attrs, str) else list(attrs)
    r = subprocess.run([exe, f"--query-gpu={','.join(attrs)}", "--format=csv,noheader,nounits", "-i", str(device_id)], capture_output=True, text=True)
    if r.returncode:
        return []
    out = []
    for tok in r.stdout.strip().split(","):
        try:
            out.append(dtype(tok.strip()))
        except ValueError:
            out.append(tok.strip())
    return out


def get_device_name(device_id: int = 0) -> str:
    return torch.cuda.get_device_name(device_id) if torch.cuda.is_available() else "cpu"


def get_max_gpu_clock_rate_in_khz(device_id: int = 0) -> int:
    nv, h = _handle(device_id)
    if h is not None:
        try:
            return int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)) * 1000
        except Exception:
            pass
    return int(get_device_info(device_id).get("clock_khz", 0))


def get_current_gpu_clock_rate_in_khz(device_id=None) -> int:
    nv, h = _handle(device_id)
    if h is not None:
        try:
            return int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)) * 1000
        except Exception:
            pass
    return 0


def is_gpu_max_performance_mode(device_id: int = 0) -> bool:
    """P0 performance state (the reference warns before benchmarking otherwise).  Unknown -> True: never block on a missing query."""
    nv, h = _handle(device_id)
    if h is not None:
        try:
            return int(nv.nvmlDeviceGetPerformanceState(h)) == 0
        except Exception:
            pass
    return True


def get_physical_device_count() -> int:
    nv = _nvml()
    if nv is not None:
        try:
            return int(nv.nvmlDeviceGetCount())
        except Exception:
            pass
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def gpu_uuid_string(uuid_bytes) -> str:
    """16 raw bytes (``cudaDeviceProp.uuid``) -> ``GPU-xxxxxxxx-xxxx-xxxx-xxxx-xxxxxxxxxxxx``."""
    h = bytes(uuid_bytes).hex()
    return f"GPU-{h[0:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:32]}"


def get_physical_gpu_uuid(gpu_index: int = 0) -> str:
    nv, h = _handle(gpu_index)
    if h is not None:
        try:
            u = nv.nvmlDeviceGetUUID(h)
            return u.decode() if isinstance(u, bytes) else str(u)
        except Exception:
            pass
    return ""


def get_nvlink_adjacency_matrix():
    """[n, n] int matrix: number of active NVLinks of GPU i whose remote end is GPU j (through an NVSwitch every link reaches the
    switch, not a peer: the matrix is then the per-GPU link count on the off-diagonal -- every peer is one hop away)."""
    n = get_physical_device_count()
    mat = [[0] * n for _ in range(n)]
    nv = _nvml()
    if nv is None or n == 0:
        return mat
    for i in range(n):
        try:
            h = nv.nvmlDeviceGetHandleByIndex(i)
        except Exception:
            continue
        active = 0
        for link in range(18):
            try:
                if nv.nvmlDeviceGetNvLinkState(h, link):
                    active += 1
            except Exception:
                break
        for j in range(n):
            if j != i:
                mat[i][j] = active
    return mat


def has_fullmesh_nvlink_pynvml() -> bool:
    m = get_nvlink_adjacency_matrix()
    return bool(m) and all(m[i][j] > 0 for i in range(len(m)) for j in range(len(m)) if i != j)


def calculate_pcie_bandwidth_gbps(generation: int, lanes: int):
    """(raw Gb/s, effective GB/s per direction) of a PCIe link: 2.5 / 5 / 8 / 16 / 32 / 64 GT/s per lane for generations 1..6,
    8b/10b encoding up to gen 2, 128b/130b from gen 3, 242B/256B FLIT mode for gen 6."""
    gts = {1: 2.5, 2: 5.0, 3: 8.0, 4: 16.0, 5: 32.0, 6: 64.0}[int(generation)]
    eff = 0.8 if generation <= 2 else (128.0 / 130.0 if generation <= 5 else 242.0 / 256.0)
    raw = gts * lanes
    return raw, raw * eff / 8.0


def get_pcie_link_max_speed_gbps(gpu_index: int = 0) -> float:
    nv, h = _handle(gpu_index)
    if h is not None:
        try:
            return calculate_pcie_bandwidth_gbps(int(nv.nvmlDeviceGetMaxPcieLinkGeneration(h)), int(nv.nvmlDeviceGetMaxPcieLinkWidth(h)))[1]
        except Exception:
            pass
    return calculate_pcie_bandwidth_gbps(5, 16)[1]            # what a B200 baseboard provides


def get_nvcc() -> str:
    from .. import _build
    return _build._nvcc()


def get_nvlink() -> str:
    """Path of the ``nvlink`` device linker next to nvcc (the reference links NVSHMEM's device library with it; nothing here needs it)."""
    import shutil
    cand = os.path.join(os.path.dirname(get_nvcc()), "nvlink")
    return cand if os.path.exists(cand) else (shutil.which("nvlink") or "")

