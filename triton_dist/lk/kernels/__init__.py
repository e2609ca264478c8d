"""Kernels written in the DSL: small SIMT examples (CPU-interpretable), a symmetric-heap ring exchange, and the tcgen05 GEMM ladder."""
