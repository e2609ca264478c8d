"""triton_dist (B200-native): compute-communication overlap library for sm_100a.

Same Python API surface as ByteDance-Seed/Triton-distributed's ``triton_dist`` package, but every fused op
is a hand-written CUDA kernel (tcgen05 / TMEM / TMA, P2P + NVLS over NVLink 5) loaded from the in-tree
``libtd_b200.so``.  No Triton, no MLIR, no NVSHMEM.  See DESIGN.md.
"""
__version__ = "0.1.0"

from . import utils  # noqa: F401
from . import _module_map as _mm

_mm.install()       # reference module paths (triton_dist.kernels.nvidia.allgather_gemm, ...) resolve to this package's modules
