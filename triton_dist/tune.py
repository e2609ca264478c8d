"""Function-level autotuner with a persistent cache and distributed (max-over-ranks) timing.

Reference: /root/reference/python/triton_dist/tune.py:280-503 -- ``@autotune(config_space, key_fn, prune_fn)``;
cache under ``~/.triton_dist/autotune/<fn>/<sha>.json`` keyed by (source hash, hardware hash, key); every config is
timed with warm-up + repetitions, and in distributed mode the time is the MAX over ranks so all ranks pick the same
winner (:485-491).  ``autotune=False`` at the call site uses the first config; ``TRITON_DIST_AUTOTUNE_ALWAYS_TUNE``
forces re-tuning.
"""
from __future__ import annotations

import functools
import hashlib
import inspect
import json
import logging
import os
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

log = logging.getLogger("triton_dist.tune")
CACHE_DIR = Path(os.environ.get("TRITON_DIST_AUTOTUNE_CACHE", str(Path.home() / ".triton_dist" / "autotune")))


def _always_tune() -> bool:
    return os.environ.get("TRITON_DIST_AUTOTUNE_ALWAYS_TUNE", "0").lower() in ("1", "true", "on")


def _hw_hash() -> str:
    if torch.cuda.is_available():
        p = torch.cuda.get_device_properties(0)
        return hashlib.sha1(f"{p.name}-{p.multi_processor_count}-{p.total_memory}-{torch.version.cuda}".encode()).hexdigest()[:12]
    return "cpu"


def _src_hash(fn: Callable) -> str:
    try:
        src = inspect.getsource(fn)
    except (OSError, TypeError):
        src = fn.__qualname__
    return hashlib.sha1(src.encode()).hexdigest()[:12]


def _time_call(thunk: Callable[[], Any], warmup: int, rep: int, pg) -> float:
    """ms per call; CUDA events on GPU, perf_counter otherwise; MAX over ranks when ``pg`` is given."""
    use_cuda = torch.cuda.is_available()
    for _ in range(warmup):
        thunk()
    if use_cuda:
        torch.cuda.synchronize()
    if pg is not None and dist.is_initialized():
        dist.barrier(group=pg)
    if use_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rep):
            thunk()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / rep
    else:
        import time
        t0 = time.perf_counter()
        for _ in range(rep):
            thunk()
        ms = (time.perf_counter() - t0) * 1e3 / rep
    if pg is not None and dist.is_initialized() and dist.get_world_size(pg) > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda" if use_cuda and dist.get_backend(pg) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=pg)
        ms = float(t.item())
    return ms


class AutoTuner:
    def __init__(self, fn: Callable, config_space: List[Dict[str, Any]], key_fn: Optional[Callable] = None,
                 prune_fn: Optional[Callable] = None, warmup: int = 5, rep: int = 10, config_kw: str = "config"):
        self.fn, self.config_space = fn, list(config_space)
        self.key_fn, self.prune_fn = key_fn, prune_fn
        self.warmup, self.rep, self.config_kw = warmup, rep, config_kw
        self.cache: Dict[str, int] = {}
        self.name = fn.__qualname__
        self._file = CACHE_DIR / self.name.replace("<", "_").replace(">", "_") / f"{_src_hash(fn)}-{_hw_hash()}.json"
        self._load()
        functools.update_wrapper(self, fn)

    def _load(self):
        try:
            self.cache = json.loads(self._file.read_text())
        except Exception:
            self.cache = {}

    def _save(self):
        try:
            self._file.parent.mkdir(parents=True, exist_ok=True)
            self._file.write_text(json.dumps(self.cache, indent=1))
        except OSError:
            pass

    def _key(self, args, kwargs) -> str:
        if self.key_fn is not None:
            return str(self.key_fn(*args, **kwargs))
        parts = []
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor):
                parts.append(f"{tuple(a.shape)}:{a.dtype}")
            elif isinstance(a, (int, float, str, bool)):
                parts.append(str(a))
        return "|".join(parts)

    def __call__(self, *args, autotune: bool = True, autotune_pg=None, **kwargs):
        space = self.config_space
        if self.prune_fn is not None:
            space = [c for c in space if self.prune_fn(c, *args, **kwargs)] or space
        if not autotune:
            return self.fn(*args, **{**kwargs, self.config_kw: space[0]})
        key = self._key(args, kwargs)
        idx = None if _always_tune() else self.cache.get(key)
        if idx is None or idx >= len(self.config_space):
            times = []
            for i, cfg in enumerate(space):
                try:
                    ms = _time_call(lambda: self.fn(*args, **{**kwargs, self.config_kw: cfg}), self.warmup, self.rep, autotune_pg)
                except Exception as e:      # an invalid config for this shape: skip, but keep ranks in lock-step
                    log.debug("config %s failed: %s", cfg, e)
                    ms = float("inf")
                times.append(ms)
            best = min(range(len(space)), key=lambda i: times[i])
            if autotune_pg is not None and dist.is_initialized() and dist.get_world_size(autotune_pg) > 1:
                obj = [best]
                dist.broadcast_object_list(obj, src=dist.get_global_rank(autotune_pg, 0), group=autotune_pg)
                best = obj[0]
            idx = self.config_space.index(space[best])
            self.cache[key] = idx
            self._save()
            log.info("autotune %s key=%s -> %s (%.3f ms)", self.name, key, self.config_space[idx], times[best])
        return self.fn(*args, **{**kwargs, self.config_kw: self.config_space[idx]})

    def best_config(self, *args, **kwargs):
        idx = self.cache.get(self._key(args, kwargs))
        return None if idx is None else self.config_space[idx]


def autotune(config_space: Iterable[Dict[str, Any]], key_fn: Optional[Callable] = None, prune_fn: Optional[Callable] = None,
             warmup: int = 5, rep: int = 10, config_kw: str = "config"):
    """Decorator: the wrapped function must accept the chosen config through keyword ``config_kw``."""
    def deco(fn):
        return AutoTuner(fn, list(config_space), key_fn, prune_fn, warmup, rep, config_kw)
    return deco


# ------------------------------------------------------------------------------------------------------------
# record keeping helpers of the reference's tune.py (hashable keys, JSON encoding of configs, hardware / version info, log control)
# ------------------------------------------------------------------------------------------------------------
def to_hashable(obj):
    """Nested lists / dicts / tensors-shapes -> a hashable, order-independent key."""
    if isinstance(obj, torch.Tensor):
        return ("tensor", tuple(obj.shape), str(obj.dtype))
    if isinstance(obj, dict):
        return tuple(sorted((k, to_hashable(v)) for k, v in obj.items()))
    if isinstance(obj, (list, tuple, set)):
        return tuple(to_hashable(v) for v in obj)
    if hasattr(obj, "__dataclass_fields__"):
        return (type(obj).__name__,) + tuple((f, to_hashable(getattr(obj, f))) for f in obj.__dataclass_fields__)
    return obj


class TuneRecordEncoder(json.JSONEncoder):
    """Serialises what tuning records contain: dataclass configs, dtypes, devices, tensors (as shape / dtype), paths."""

    def default(self, o):
        if hasattr(o, "__dataclass_fields__"):
            return {"__dataclass__": type(o).__name__, **{f: getattr(o, f) for f in o.__dataclass_fields__}}
        if isinstance(o, (torch.dtype, torch.device, Path)):
            return str(o)
        if isinstance(o, torch.Tensor):
            return {"__tensor__": list(o.shape), "dtype": str(o.dtype)}
        return super().default(o)


def from_json(text: str):
    return json.loads(text)


def store_autotune_data(path, data: Dict[str, Any]):
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(json.dumps(data, indent=1, cls=TuneRecordEncoder))


def load_autotune_data(path) -> Dict[str, Any]:
    try:
        return json.loads(Path(path).read_text())
    except (OSError, ValueError):
        return {}


def pretty_triton_config_repr(cfg) -> str:
    """One-line description of a configuration (dataclass, dict, or anything with a repr)."""
    if hasattr(cfg, "__dataclass_fields__"):
        return type(cfg).__name__ + "(" + ", ".join(f"{f}={getattr(cfg, f)}" for f in cfg.__dataclass_fields__) + ")"
    if isinstance(cfg, dict):
        return ", ".join(f"{k}={v}" for k, v in cfg.items())
    return repr(cfg)


def get_hardware_info() -> Dict[str, Any]:
    if not torch.cuda.is_available():
        return {"device": "cpu"}
    p = torch.cuda.get_device_properties(0)
    return {"device": p.name, "sms": p.multi_processor_count, "memory_gb": round(p.total_memory / 2 ** 30, 1), "cuda": torch.version.cuda,
            "capability": f"{p.major}.{p.minor}", "num_gpus": torch.cuda.device_count()}


hw_hash = _hw_hash


def get_triton_dist_version() -> str:
    from . import __version__
    return __version__


def get_git_info() -> Dict[str, str]:
    import subprocess
    root = Path(__file__).resolve().parent.parent
    try:
        rev = subprocess.run(["git", "-C", str(root), "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip()
        dirty = bool(subprocess.run(["git", "-C", str(root), "status", "--porcelain"], capture_output=True, text=True, timeout=5).stdout.strip())
    except Exception:      # noqa: BLE001
        rev, dirty = "", False
    return {"commit": rev, "dirty": str(dirty)}


def get_deps() -> Dict[str, str]:
    return {"torch": torch.__version__, "cuda": str(torch.version.cuda), "triton_dist": get_triton_dist_version()}


def get_cuda_extra_args() -> Dict[str, Any]:
    """What a cache entry must also depend on for CUDA kernels (the reference hashes Triton launch extras): the native library digest."""
    from . import _build
    lib = _build.LIBDIR / "libtd_b200.so"
    return {"lib_mtime": int(lib.stat().st_mtime) if lib.exists() else 0}


def log_to_file(path: str, level: int = logging.INFO):
    h = logging.FileHandler(path)
    h.setLevel(level)
    h.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s"))
    log.addHandler(h)
    log.setLevel(min(log.level or level, level))
    return h


def set_stream_handler_log_level(level: int):
    found = False
    for h in log.handlers:
        if isinstance(h, logging.StreamHandler) and not isinstance(h, logging.FileHandler):
            h.setLevel(level)
            found = True
    if not found:
        h = logging.StreamHandler()
        h.setLevel(level)
        log.addHandler(h)
    log.setLevel(min(log.level or level, level))
