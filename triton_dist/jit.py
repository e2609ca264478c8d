"""Compile-and-load for user CUDA kernels written against the device header (``csrc/td/primitives.cuh`` / ``ptx.cuh``).

The reference lets users write distributed kernels in its Triton dialect (``triton_dist.jit``, python/triton_dist/jit.py:274-313) or in
the little_kernel DSL whose runtime shells out to nvcc and loads the result through the driver API
(python/little_kernel/runtime/compiler.py, cuda_runtime.py).  Here a kernel is CUDA C++: ``compile_cuda(source)`` runs
``nvcc -gencode arch=compute_100a,code=sm_100a`` with the header tree on the include path, caches the shared object by content hash and
returns a ``ctypes.CDLL``.  The source provides its own ``extern "C"`` launcher(s); ``symm_args()`` gives the struct the device
primitives need (rank, world, heap base / stride, multicast base).

    lib = compile_cuda(r'''
        #include "td/primitives.cuh"
        using namespace td;
        __global__ void ring(SymmCtx c, uint32_t* flag, float* data, uint32_t round) {
          const int nxt = (c.rank + 1) % c.world;
          symm_at(c, data, nxt)[threadIdx.x] = c.rank * 100.f + round;           // data, then ...
          __syncthreads();
          if (threadIdx.x == 0) notify(c, flag, nxt, round);                      // ... the flag (release, system scope)
          if (threadIdx.x < 32) wait<true, true>(flag, 1, round);                 // my predecessor's flag (acquire)
        }
        extern "C" void launch_ring(SymmCtx c, void* flag, void* data, unsigned round, void* stream) {
          ring<<<1, 64, 0, (cudaStream_t)stream>>>(c, (uint32_t*)flag, (float*)data, round);
        }''')
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path
from typing import Sequence

from . import _build

_CACHE = Path(os.environ.get("TD_JIT_CACHE", str(_build.ROOT / "build" / "jit")))


class SymmCtx(C.Structure):
    """Mirror of ``td::SymmCtx`` (csrc/td/primitives.cuh) for by-value kernel arguments."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("base", C.c_ulonglong), ("stride", C.c_ulonglong), ("mc_base", C.c_ulonglong)]


def symm_ctx() -> SymmCtx:
    """The symmetric-heap description of this process (after ``initialize_distributed``)."""
    from . import utils as U
    r, w, base, stride, mc = U.symm_ctx_fields()
    return SymmCtx(int(r), int(w), int(base), int(stride), int(mc))


_HDR_DIGEST = None


def _headers_digest() -> str:
    global _HDR_DIGEST
    if _HDR_DIGEST is None:
        h = hashlib.sha256()
        for f in sorted((_build.CSRC / "td").glob("*.cuh")):
            h.update(f.read_bytes())
        _HDR_DIGEST = h.hexdigest()[:16]
    return _HDR_DIGEST


def compile_cuda(source: str, extra_flags: Sequence[str] = (), name: str = "kernel") -> C.CDLL:
    """nvcc -> shared object -> ``ctypes.CDLL``.  Works without a GPU (cross-compiles sm_100a); launching needs one."""
    flags = list(_build.GENCODE) + ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-cudart", "shared",
                                    "-I", str(_build.CSRC)] + list(extra_flags)
    tmo = _build.debug_wait_timeout_ns()
    if tmo is not None:               # TD_DEBUG_WAITS: user / DSL kernels get the guarded waits too
        flags.append(f"-DTD_WAIT_TIMEOUT_NS={tmo}ull")
    # the key must not depend on where the checkout lives (the cache travels with the tree to the GPU box), but it must change when
    # the device headers do
    key = hashlib.sha256((source + "\0" + " ".join(flags).replace(str(_build.CSRC), "$CSRC") + "\0" + _headers_digest()).encode()).hexdigest()[:16]
    _CACHE.mkdir(parents=True, exist_ok=True)
    so = _CACHE / f"{name}_{key}.so"
    if not so.exists():
        src = _CACHE / f"{name}_{key}.cu"
        src.write_text(source)
        tmp = _CACHE / f"{name}_{key}.{os.getpid()}.tmp.so"
        r = subprocess.run([_build._nvcc(), "-shared", *flags, "-o", str(tmp), str(src)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{r.stderr[-4000:]}")
        os.replace(tmp, so)
    return C.CDLL(str(so))


# ------------------------------------------------------------------------------------------------------------
# launch helpers (reference: little_kernel/runtime/{kernel,tma_descriptor}.py -- argument marshalling, cluster dims, TMA maps)
# ------------------------------------------------------------------------------------------------------------
def _marshal(arg):
    """torch tensors -> device pointers, ints / floats / bools -> C scalars, SymmCtx / ctypes values pass through."""
    import torch
    if isinstance(arg, torch.Tensor):
        return C.c_void_p(arg.data_ptr())
    if isinstance(arg, bool):
        return C.c_int(int(arg))
    if isinstance(arg, int):
        return C.c_longlong(arg) if abs(arg) >= 2 ** 31 else C.c_int(arg)
    if isinstance(arg, float):
        return C.c_float(arg)
    if arg is None:
        return C.c_void_p(None)
    return arg


class JitKernel:
    """A launcher exported by a ``compile_cuda`` module, callable with tensors / scalars.  The C launcher's last parameter must be
    the ``void* stream``; ``kernel(*args, stream=None)`` appends the current CUDA stream."""

    def __init__(self, lib: C.CDLL, name: str):
        self.fn = getattr(lib, name)
        self.fn.restype = None
        self.name = name

    def __call__(self, *args, stream=None):
        import torch
        s = (stream or torch.cuda.current_stream()).cuda_stream if torch.cuda.is_available() else 0
        self.fn(*[_marshal(a) for a in args], C.c_void_p(s))


def kernel(source: str, launcher: str, extra_flags: Sequence[str] = ()) -> JitKernel:
    """``compile_cuda`` + argument marshalling: returns a callable for the ``extern "C"`` launcher ``launcher``."""
    return JitKernel(compile_cuda(source, extra_flags, name=launcher), launcher)


def make_tma_2d(t, box_inner: int, box_outer: int, swizzle: int = 128):
    """A ``CUtensorMap`` (128-byte opaque ctypes buffer) over a 2-D row-major tensor -- what little_kernel's
    ``create_tma_2d_descriptor`` returns; pass it by value (``__grid_constant__ const CUtensorMap``) to a JIT kernel."""
    from . import _C
    lib = _C.cuda_lib()
    if not hasattr(lib, "td_make_tma_2d"):
        raise RuntimeError("td_make_tma_2d is not exported by libtd_b200.so")
    buf = (C.c_ubyte * 128)()
    fn = lib.td_make_tma_2d
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]
    rows, cols = t.shape
    _C.check(fn(buf, t.data_ptr(), rows, cols, t.stride(0), t.element_size(), box_inner, box_outer, swizzle), "td_make_tma_2d")
    return buf


def jit(fn=None, **kernel_options):
    """``@triton_dist.jit.jit`` -- the decorator role of the reference's ``triton_dist.jit`` (jit.py:274-313: ``triton.jit`` plus the
    NVSHMEM device library).  Kernels are authored in the Python DSL here: this is :func:`triton_dist.lk.kernel` (typed AST -> CUDA C++
    over csrc/td/*.cuh -> nvcc), and the symmetric-heap device API comes with the generated source's headers instead of a bitcode link."""
    from . import lk
    return lk.kernel(fn, **kernel_options) if fn is not None else lk.kernel(**kernel_options)

