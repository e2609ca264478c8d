"""Runtime: process bring-up, symmetric tensors, barriers, env helpers, topology.

API-compatible with the reference's ``triton_dist.utils`` (/root/reference/python/triton_dist/utils.py) --
``initialize_distributed``, ``nvshmem_create_tensor(s)``, ``nvshmem_free_tensor_sync``,
``nvshmem_barrier_all_on_stream``, ``NVSHMEM_SIGNAL_DTYPE``, ``dist_print``, ``get_bool_env`` ... -- but backed by
our own CUDA-VMM symmetric heap (or the shared-memory emulation when there is no GPU) instead of NVSHMEM.
"""
from __future__ import annotations

import ctypes as C
import datetime
import os
import random
import sys
import time
from contextlib import contextmanager
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .. import _C
from .symm import SymmetricHeap, tensor_from_ptr  # noqa: F401

# signals are 32-bit words in the symmetric heap (the reference uses uint64 because NVSHMEM does: utils.py:572)
NVSHMEM_SIGNAL_DTYPE = torch.int32
SIGNAL_DTYPE = torch.int32

_STATE = {
    "initialized": False, "rank": 0, "world": 1, "local_rank": 0, "local_world": 1, "group": None,
    "heap": None, "device": torch.device("cpu"), "barrier_ctx": None,
}


# ------------------------------------------------------------------------------------------------------------
# env helpers (utils.py:890-911)
# ------------------------------------------------------------------------------------------------------------
def accept_ref_hints(fn: str, kwargs: dict, allowed: tuple = ()) -> None:
    """Entry points keep the reference's signatures; arguments that only steer the reference's Triton kernels or its host
    streams (tile hints, side streams) are accepted by NAME and documented as having no effect here.  Anything else raises
    -- an argument is never silently dropped."""
    bad = [k for k in kwargs if k not in allowed]
    if bad:
        raise TypeError(f"{fn}() got unexpected keyword argument(s) {bad}; reference-only hints accepted (no effect on "
                        f"the sm_100a kernels): {list(allowed)}")
    for k in ("A_scale", "B_scale", "As", "Bs", "scale_a", "scale_b"):
        if kwargs.get(k) is not None:
            raise NotImplementedError(f"{fn}(): {k} was given but this entry point does not apply scales")


def get_bool_env(name: str, default: bool = False) -> bool:
    v = os.environ.get(name)
    if v is None:
        return default
    return v.strip().lower() in ("1", "true", "on", "yes", "y")


def get_int_env(name: str, default: int) -> int:
    v = os.environ.get(name)
    return default if v is None or v == "" else int(v)


def _parse_size(s: str) -> int:
    s = s.strip().lower()
    mult = 1
    for suf, m in (("k", 1 << 10), ("m", 1 << 20), ("g", 1 << 30)):
        if s.endswith(suf) or s.endswith(suf + "b"):
            mult = m
            s = s.rstrip("b").rstrip(suf)
            break
    return int(float(s) * mult)


def backend() -> str:
    """'cuda' on a GPU box, 'host' (shared-memory emulation) otherwise."""
    return "cuda" if torch.cuda.is_available() and not get_bool_env("TD_FORCE_HOST_BACKEND") else "host"


# ------------------------------------------------------------------------------------------------------------
# bring-up (utils.py:335-367)
# ------------------------------------------------------------------------------------------------------------
def init_seed(seed: int = 0):
    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":16:8")
    torch.manual_seed(seed)
    np.random.seed(seed % (2 ** 31))
    random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False


def initialize_distributed(seed: Optional[int] = None, initialize_shmem: bool = True,
                           heap_bytes: Optional[int] = None):
    """Create the process group (NCCL+gloo on GPU boxes, gloo otherwise) and the symmetric heap.

    Reads RANK / LOCAL_RANK / WORLD_SIZE / LOCAL_WORLD_SIZE (torchrun).  Works single-process too (world=1
    without torchrun).  Returns the tensor-parallel process group, like the reference.
    """
    if _STATE["initialized"]:
        return _STATE["group"]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    use_cuda = backend() == "cuda"
    if use_cuda:
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    else:
        device = torch.device("cpu")
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        os.environ.setdefault("RANK", str(rank))
        os.environ.setdefault("WORLD_SIZE", str(world))
        kwargs = dict(world_size=world, rank=rank, timeout=datetime.timedelta(seconds=1800))
        if use_cuda:
            dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=device, **kwargs)
        else:
            dist.init_process_group(backend="gloo", **kwargs)
    group = dist.new_group(ranks=list(range(world)), backend="nccl" if use_cuda else "gloo")
    _STATE.update(initialized=True, rank=rank, world=world, local_rank=local_rank, local_world=local_world,
                  group=group, device=device)
    init_seed(seed if seed is not None else rank)
    if use_cuda:
        torch.cuda.synchronize()
    dist.barrier(group=_gloo_group())
    if initialize_shmem:
        init_symmetric_heap(heap_bytes)
    return group


_GLOO = {"g": None}


def _gloo_group():
    """A gloo group for host-side object collectives (fd exchange, barriers that must not touch the GPU)."""
    if not dist.is_initialized():
        return None
    if _GLOO["g"] is None:
        if dist.get_backend() == "gloo":
            _GLOO["g"] = dist.group.WORLD
        else:
            _GLOO["g"] = dist.new_group(backend="gloo")
    return _GLOO["g"]


def init_symmetric_heap(heap_bytes: Optional[int] = None) -> SymmetricHeap:
    if _STATE["heap"] is not None:
        return _STATE["heap"]
    if heap_bytes is None:
        env = os.environ.get("TD_SYMM_HEAP_SIZE") or os.environ.get("NVSHMEM_SYMMETRIC_SIZE")
        heap_bytes = _parse_size(env) if env else ((4 << 30) if backend() == "cuda" else (64 << 20))
    heap = SymmetricHeap(heap_bytes, _STATE["rank"], _STATE["world"], _STATE["device"], group=_gloo_group())
    _STATE["heap"] = heap
    return heap


def finalize_distributed():
    heap = _STATE["heap"]
    if heap is not None:
        if dist.is_initialized():
            if _STATE["device"].type == "cuda":
                torch.cuda.synchronize()
            dist.barrier(group=_gloo_group())
        _STATE["barrier_ctx"] = None
        heap.destroy()
        _STATE["heap"] = None
    if dist.is_initialized():
        dist.destroy_process_group()
    _GLOO["g"] = None
    _STATE.update(initialized=False, group=None)


def get_heap() -> SymmetricHeap:
    if _STATE["heap"] is None:
        if not _STATE["initialized"]:
            initialize_distributed()
        else:
            init_symmetric_heap()
    return _STATE["heap"]


def rank() -> int:
    return _STATE["rank"]


def world_size() -> int:
    return _STATE["world"]


def get_triton_dist_world():
    return _STATE["group"]


def get_triton_dist_local_world_size() -> int:
    return _STATE["local_world"]


def current_device() -> torch.device:
    return _STATE["device"]


# ------------------------------------------------------------------------------------------------------------
# symmetric tensors (utils.py:246-287)
# ------------------------------------------------------------------------------------------------------------
def nvshmem_create_tensor(shape, dtype: torch.dtype) -> torch.Tensor:
    """Collective: allocate a zero-filled tensor at the same heap offset on every rank."""
    return get_heap().tensor(shape, dtype)


def nvshmem_create_tensors(shape, dtype: torch.dtype, rank: int, local_world_size: int) -> List[torch.Tensor]:
    """Collective: allocate a symmetric tensor and return views of it on every local rank (index = rank)."""
    heap = get_heap()
    t = heap.tensor(shape, dtype)
    return [heap.peer_view(t, r) for r in range(local_world_size)]


def nvshmem_free_tensor_sync(t: torch.Tensor):
    """Collective free (like the reference's ``nvshmem_free``): drain my stream, wait until every rank has done the same,
    then release the offset -- a peer may still be writing flags / data into the tensor until it reaches the barrier."""
    heap = get_heap()
    if heap.is_cuda:
        torch.cuda.synchronize()
    barrier_all_host()
    heap.free_tensor(t)


symm_tensor = nvshmem_create_tensor
symm_tensors = nvshmem_create_tensors
symm_free = nvshmem_free_tensor_sync


def symm_at(t: torch.Tensor, peer: int) -> torch.Tensor:
    """Host-side mirror of the device primitive: the peer's copy of a symmetric tensor."""
    return get_heap().peer_view(t, peer)


def symm_ctx_fields():
    """(rank, world, base, stride, mc_base) -- what every distributed kernel receives as ``td::SymmCtx``."""
    h = get_heap()
    return h.rank, h.world, h.base, h.stride, h.mc_base


def is_nvshmem_multimem_supported() -> bool:
    """True when the NVLS multicast mapping of the heap exists (reference: utils.py:756-792)."""
    return bool(_STATE["heap"] is not None and _STATE["heap"].mc_base)


def has_tma() -> bool:
    return torch.cuda.is_available() and torch.cuda.get_device_capability()[0] >= 9


def supports_p2p_native_atomic() -> bool:
    if backend() != "cuda" or _STATE["world"] == 1:
        return True
    lib = _C.cuda_lib()
    me = _STATE["local_rank"]
    return all(lib.td_p2p_native_atomics(me, p) for p in range(_STATE["local_world"]) if p != me)


# ------------------------------------------------------------------------------------------------------------
# barriers
# ------------------------------------------------------------------------------------------------------------
class BarrierAllContext:
    """Flag-flip cross-rank barrier state living in the symmetric heap (common_ops.py:227-261).

    The epoch counter is device resident, so ``barrier_all_on_stream`` can be captured in a CUDA graph
    (the reference's cannot: common_ops.py:251-253)."""

    def __init__(self, is_intra_node: bool = True):
        heap = get_heap()
        self.heap = heap
        self.slots = heap.tensor((2 * max(heap.world, 1),), torch.int32)
        self.epoch = torch.zeros(4, dtype=torch.int32, device=heap.device)   # local, not symmetric
        self.host_epoch = 0
        barrier_all_host()


def barrier_all_host():
    """Host-level rendezvous (gloo): used around collective allocations, never on a hot path."""
    if _STATE["device"].type == "cuda":
        torch.cuda.synchronize()
    if dist.is_initialized() and _STATE["world"] > 1:
        dist.barrier(group=_gloo_group())


def barrier_all_on_stream(ctx: Optional[BarrierAllContext] = None, stream=None):
    """Cross-rank barrier ordered on ``stream`` (device kernel on GPU, atomic flag-flip on the host backend)."""
    if ctx is None:
        if _STATE["barrier_ctx"] is None:
            _STATE["barrier_ctx"] = BarrierAllContext()
        ctx = _STATE["barrier_ctx"]
    heap = ctx.heap
    if heap.is_cuda:
        from ..ops import comm
        comm.barrier_all(ctx, stream)
    else:
        ctx.host_epoch += 1
        lib = _C.host_lib()
        rc = lib.tdh_barrier_all(heap._handle, heap.offset_of(ctx.slots), ctx.host_epoch,
                                 get_int_env("TD_HOST_TIMEOUT_US", 60_000_000))
        if rc:
            raise TimeoutError("barrier_all timed out: " + lib.tdh_last_error().decode())


def nvshmem_barrier_all_on_stream(stream=None):
    barrier_all_on_stream(None, stream)


# ------------------------------------------------------------------------------------------------------------
# stream-ordered flag ops (copy-engine style; common_ops.py:364-414)
# ------------------------------------------------------------------------------------------------------------
def _stream_ptr(stream) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def set_signal(signal_ptr_or_tensor, value: int, stream=None):
    ptr = signal_ptr_or_tensor.data_ptr() if isinstance(signal_ptr_or_tensor, torch.Tensor) else int(signal_ptr_or_tensor)
    if backend() == "cuda":
        _C.check(_C.cuda_lib().td_stream_write_value32(C.c_void_p(_stream_ptr(stream)), ptr, value & 0xFFFFFFFF),
                 "cuStreamWriteValue32")
    else:
        _C.host_lib().tdh_notify32(C.c_void_p(ptr), value & 0xFFFFFFFF, 1)


def wait_eq(signal_ptr_or_tensor, value: int, stream=None, geq: bool = False):
    ptr = signal_ptr_or_tensor.data_ptr() if isinstance(signal_ptr_or_tensor, torch.Tensor) else int(signal_ptr_or_tensor)
    if backend() == "cuda":
        _C.check(_C.cuda_lib().td_stream_wait_value32(C.c_void_p(_stream_ptr(stream)), ptr, value & 0xFFFFFFFF, int(geq)),
                 "cuStreamWaitValue32")
    else:
        rc = _C.host_lib().tdh_wait32(C.c_void_p(ptr), value & 0xFFFFFFFF, 1 if geq else 0,
                                      get_int_env("TD_HOST_TIMEOUT_US", 60_000_000))
        if rc:
            raise TimeoutError(f"wait on signal {ptr:#x} == {value} timed out")


# ------------------------------------------------------------------------------------------------------------
# misc helpers (utils.py:396-470, 814-1045)
# ------------------------------------------------------------------------------------------------------------
def dist_print(*args, allowed_ranks="all", prefix: bool = False, need_sync: bool = False, **kwargs):
    r, w = _STATE["rank"], _STATE["world"]
    if allowed_ranks == "all":
        allowed_ranks = list(range(w))
    if need_sync and dist.is_initialized():
        for i in range(w):
            if i == r and r in allowed_ranks:
                print(*(([f"[rank:{r}]"] if prefix else []) + list(args)), **kwargs)
                sys.stdout.flush()
            dist.barrier(group=_gloo_group())
        return
    if r in allowed_ranks:
        print(*(([f"[rank:{r}]"] if prefix else []) + list(args)), **kwargs)
        sys.stdout.flush()


def rand_tensor(shape, dtype: torch.dtype, device=None, scale: float = 1.0) -> torch.Tensor:
    device = device or _STATE["device"]
    if dtype in (torch.int8, torch.int32, torch.int64, torch.uint8):
        return torch.randint(-8, 8, shape, device=device).to(dtype)
    t = (torch.rand(shape, device=device, dtype=torch.float32) * 2 - 1) * scale
    return t.to(dtype)


def sleep_async(ms: float):
    """Queue ~ms of GPU idle time on the current stream (straggler injection; reference: utils.py sleep_async)."""
    if torch.cuda.is_available():
        clock_khz = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1_500_000
        torch.cuda._sleep(int(ms * clock_khz))
    else:
        time.sleep(ms / 1e3)


def launch_cooperative_grid_options():
    """Kept for API parity; our persistent kernels size their grid <= #SM so co-residency is structural."""
    return {}


def wait_until_max_gpu_clock_or_warning(*_a, **_k):
    return True


@contextmanager
def with_torch_deterministic(mode: bool = True):
    old = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(mode, warn_only=True)
    try:
        yield
    finally:
        torch.use_deterministic_algorithms(old, warn_only=True)


def barrier_async(pg=None):
    """One-int all-reduce used as a cheap ordered rendezvous by the autotuner (utils.py:935)."""
    if dist.is_initialized() and _STATE["world"] > 1:
        t = torch.ones(1, dtype=torch.int32, device=_STATE["device"])
        dist.all_reduce(t, group=pg or _STATE["group"])


from .lazy import LazyAllocator, LazyTensor, NVSHMEMLazyAllocator  # noqa: E402,F401
from .topology import (get_device_info, get_intranode_max_speed_gbps, get_nvlink_max_speed_gbps,  # noqa: E402,F401
                       has_fullmesh_nvlink, get_numa_node)


def generate_data(configs, device=None):
    """Yield fresh random tensors for ``[(shape, dtype, scale), ...]`` (reference: utils.py:_make_tensor/generate_data):
    tests draw new inputs every iteration so stale-buffer bugs cannot hide behind identical data."""
    device = device or _STATE["device"]
    while True:
        yield [rand_tensor(shape, dtype, device, scale) for (shape, dtype, scale) in configs]


def _make_tensor(shape, dtype, init_args, device=None):
    """One random tensor ``scale * randn + bias`` (reference: utils.py ``_make_tensor``; ``init_args`` = ``(scale, bias)`` or ``scale``)."""
    scale, bias = init_args if isinstance(init_args, (tuple, list)) else (init_args, 0.0)
    t = rand_tensor(shape, dtype, device or _STATE["device"], scale)
    return t + bias if bias else t


def triton_dist_key() -> str:
    """Hash of the native sources (cache-key ingredient of the autotuner; reference: utils.py triton_dist_key)."""
    import hashlib
    from pathlib import Path
    h = hashlib.sha1()
    for p in sorted((Path(__file__).resolve().parents[2] / "csrc").rglob("*.cu*")):
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------------------
# reference spellings of the stream-ordered flag / copy helpers and the device search helpers
# (kernels/nvidia/common_ops.py:264-414)
# ------------------------------------------------------------------------------------------------------------
def _set_signal_cuda(signal, value: int, stream=None):
    """cuStreamWriteValue32 on ``stream`` (``signal``: 1-element int32 tensor or raw pointer)."""
    return set_signal(signal, value, stream)


def _wait_eq_cuda(signal, value: int, stream=None):
    """cuStreamWaitValue32(EQ) on ``stream``."""
    return wait_eq(signal, value, stream)


def _memcpy_async_cuda(dst: torch.Tensor, src: torch.Tensor, stream=None):
    """cudaMemcpyAsync on ``stream`` (peer views of the symmetric heap are ordinary device pointers: the copy engine moves them)."""
    if stream is None or not dst.is_cuda:
        dst.copy_(src, non_blocking=True)
        return dst
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
    return dst


def bisect_left(sorted_values: torch.Tensor, x) -> torch.Tensor:
    """First index i with sorted_values[i] >= x (the reference's device search helper; here ``torch.searchsorted``)."""
    xv = torch.as_tensor(x, device=sorted_values.device, dtype=sorted_values.dtype)
    return torch.searchsorted(sorted_values, xv, right=False)


def bisect_right(sorted_values: torch.Tensor, x) -> torch.Tensor:
    xv = torch.as_tensor(x, device=sorted_values.device, dtype=sorted_values.dtype)
    return torch.searchsorted(sorted_values, xv, right=True)


def get_device_property(device=None):
    """``torch.cuda.get_device_properties`` of the current device (SM count, name, memory); None without a GPU."""
    if not torch.cuda.is_available():
        return None
    return torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device())

from .extras import (CUDA_CHECK, TorchStreamWrapper, cuda_occupancy_max_activate_blocks_per_multiprocessor, get_cpu_info_linux,  # noqa: E402,F401
                     get_device_max_shared_memory_size, get_dtype_size, get_group_numa_world_size, get_numa_node_count_in_group,
                     get_nvshmem_hash, get_nvshmem_version, get_shmem_backend, get_shmem_hash, get_shmem_version, get_smi_device_index,
                     init_nvshmem_by_torch_process_group, is_cuda, is_fp8_dtype, is_hip, is_maca, is_mori_shmem, is_rocshmem,
                     is_shmem_initialized, requires, requires_p2p_native_atomic, support_launch_cooperative_grid,
                     torch_stream_max_priority, triton_packed_version, warn_if_cuda_launch_blocking)
from .lazy import LazyTensorSpec, get_underlying_tensor, nvshmem_free_lazy_tensor  # noqa: E402,F401


def _is_cuda_launch_blocking() -> bool:
    return os.environ.get("CUDA_LAUNCH_BLOCKING", "0") == "1"


def _torch_has_fp8() -> bool:
    return hasattr(torch, "float8_e4m3fn") and hasattr(torch, "float8_e5m2")

