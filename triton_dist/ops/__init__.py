"""Ops = host orchestration + sm_100a device kernels (the reference's ``triton_dist.kernels.nvidia``)."""
from .gemm import GemmConfig, gemm, default_config  # noqa: F401
