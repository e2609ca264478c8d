"""Small helpers of the reference's ``triton_dist/utils.py`` that scripts and tests written against it use: platform predicates,
error checks, stream wrapper, NUMA queries, decorators, dtype sizes.  Vendor-specific entry points (rocSHMEM / MORI / MACA) are out of
scope; the NVSHMEM-named ones map onto the symmetric heap of this framework."""
from __future__ import annotations

import functools
import os
import warnings
from typing import Optional

import torch


# ---- platform ---------------------------------------------------------------------------------------------------------
def is_cuda() -> bool:
    return True


def is_hip() -> bool:
    return False


def is_maca() -> bool:
    return False


def get_shmem_backend() -> str:
    """The reference answers nvshmem / rocshmem / mori; here the symmetric heap is this framework's own (CUDA VMM + NVLS)."""
    return "td_symm_heap"


def is_rocshmem() -> bool:
    return False


def is_mori_shmem() -> bool:
    return False


def get_shmem_version() -> str:
    from .. import __version__
    return __version__


def get_shmem_hash() -> str:
    """Content hash of the device runtime sources (what the reference derives from the NVSHMEM bitcode)."""
    import hashlib
    from .. import _build
    h = hashlib.sha256()
    for f in sorted((_build.CSRC / "runtime").glob("*")) + sorted((_build.CSRC / "td").glob("*.cuh")):
        if f.is_file():
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


get_nvshmem_version, get_nvshmem_hash = get_shmem_version, get_shmem_hash


def is_shmem_initialized() -> bool:
    from . import _STATE
    return _STATE.get("heap") is not None


def init_nvshmem_by_torch_process_group(pg=None):
    """The reference bootstraps NVSHMEM over a torch process group; here ``initialize_distributed`` creates the symmetric heap, so this
    only checks that it happened."""
    if not is_shmem_initialized():
        raise RuntimeError("call triton_dist.utils.initialize_distributed() first (it creates the symmetric heap)")


# ---- CUDA helpers -----------------------------------------------------------------------------------------------------
def CUDA_CHECK(err):
    """Accepts cuda-python ``CUresult`` / ``cudaError_t`` values, (err, ...) tuples, and plain integer codes."""
    if isinstance(err, tuple):
        err = err[0]
    code = int(getattr(err, "value", err))
    if code != 0:
        name = getattr(err, "name", None) or f"error {code}"
        raise RuntimeError(f"Cuda Error: {name}")


class TorchStreamWrapper:
    """``__cuda_stream__`` protocol object for cuda-python calls (reference: utils.py:308-317)."""

    def __init__(self, pt_stream: "torch.cuda.Stream"):
        self.pt_stream = pt_stream
        self.handle = pt_stream.cuda_stream

    def __cuda_stream__(self):
        return (0, self.pt_stream.cuda_stream)


def torch_stream_max_priority() -> int:
    try:
        _, high = torch.cuda.Stream.priority_range()
    except Exception:      # noqa: BLE001
        high = -1
    return high


def get_device_max_shared_memory_size(device_id: int = 0) -> int:
    if torch.cuda.is_available():
        p = torch.cuda.get_device_properties(device_id)
        return int(getattr(p, "shared_memory_per_block_optin", 227 * 1024))
    return 227 * 1024            # sm_100a


def support_launch_cooperative_grid() -> bool:
    return True                   # persistent kernels here size their grids to the SM count; no cooperative-launch flag is needed


def cuda_occupancy_max_activate_blocks_per_multiprocessor(kernel, num_warps: int = 4, *_, dynamic_smem: int = 0, **__) -> int:
    """Occupancy of a DSL kernel (``triton_dist.lk.Kernel``): resident CTAs per SM from its registers and shared memory."""
    attrs = kernel.attributes()
    threads = num_warps * 32
    smem = attrs["static_smem"] + (dynamic_smem or getattr(kernel, "dyn_smem_bytes", 0))
    by_regs = (64 * 1024) // max(1, attrs["regs"] * threads)
    by_smem = (227 * 1024) // max(1, smem) if smem else 32
    by_threads = 2048 // threads
    return max(0, min(by_regs, by_smem, by_threads, 32))


def warn_if_cuda_launch_blocking():
    if os.environ.get("CUDA_LAUNCH_BLOCKING", "0") not in ("", "0"):
        warnings.warn("CUDA_LAUNCH_BLOCKING is set: kernels that wait for other streams / ranks can deadlock and every timing is serialised")


def get_smi_device_index(device_id: int) -> int:
    """Index of a torch device in nvidia-smi order (CUDA_VISIBLE_DEVICES applied)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if device_id < len(ids) and ids[device_id].isdigit():
            return int(ids[device_id])
    return device_id


# ---- host / NUMA ------------------------------------------------------------------------------------------------------
def get_cpu_info_linux():
    vendor = model = None
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("vendor_id"):
                    vendor = line.split(":", 1)[1].strip()
                elif line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                if vendor and model:
                    break
    except OSError:
        pass
    return vendor, model


def get_numa_node_count_in_group(pg) -> int:
    """Distinct NUMA nodes of the GPUs in ``pg`` (2 only when the ranks split evenly over two nodes, else 1 -- the reference's rule)."""
    from . import topology
    import torch.distributed as dist
    n = pg.size() if pg is not None else dist.get_world_size()
    mine = topology.get_numa_node(torch.cuda.current_device()) if torch.cuda.is_available() else 0
    nodes = [None] * n
    dist.all_gather_object(nodes, mine, group=pg)
    uniq = sorted(set(nodes))
    if len(uniq) != 2 or nodes.count(uniq[0]) != nodes.count(uniq[1]):
        return 1
    return 2


def get_group_numa_world_size(pg) -> int:
    return pg.size() // get_numa_node_count_in_group(pg)


# ---- decorators -------------------------------------------------------------------------------------------------------
def requires(condition_func):
    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            assert condition_func(), f"{condition_func.__name__} is needed for {func.__name__}, please check..."
            return func(*args, **kwargs)
        return wrapper
    return decorator


def requires_p2p_native_atomic(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        from . import supports_p2p_native_atomic
        if not supports_p2p_native_atomic():
            warnings.warn(f"function {fn.__name__} requires P2P native atomics, which this platform does not report")
        return fn(*args, **kwargs)
    return wrapper


# ---- dtypes -----------------------------------------------------------------------------------------------------------
def is_fp8_dtype(dtype: torch.dtype) -> bool:
    return dtype in (torch.float8_e4m3fn, torch.float8_e5m2, torch.float8_e4m3fnuz, torch.float8_e5m2fnuz)


def get_dtype_size(dtype: torch.dtype) -> int:
    return torch.empty(0, dtype=dtype).element_size()


def triton_packed_version() -> Optional[str]:
    return None                   # no Triton in this stack
