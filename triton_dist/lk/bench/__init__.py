"""Micro-benchmarks written in the DSL: the measured denominators a kernel author compares against.

Reference: python/little_kernel/benchmark/{compute,memory,latency,warp,sm}/ + run_all.py (FLOPS / SFU / wgmma, HBM / L2 / L1 / shared /
register-file bandwidth, bank conflicts, TMA layouts, arithmetic / memory / sync latencies, shuffles and votes, IPC, occupancy).
The same questions for sm_100a, as short ``@lk.kernel`` functions (``kernels.py``) with CUDA-event runners (``run.py``):

    python -m triton_dist.lk.bench            # every benchmark, one JSON line + a table (single GPU)
    python -m triton_dist.lk.bench --only global_copy,smem_stride --json out.json

Every kernel also produces a checkable result (a copy is a copy, a pointer chase ends where the permutation says, a reduction has a
closed form), so the CPU interpreter runs the whole suite at toy sizes (tests/test_lk_cpu.py) -- the numbers need a GPU, the kernels'
correctness does not.
"""
from .kernels import KERNELS  # noqa: F401
from .run import BENCHES, run_all  # noqa: F401
