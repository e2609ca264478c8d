"""Kernels written in the DSL (all of them run in the CPU interpreter; the distributed ones across processes on the emulation backend).

* ``simt``                  SIMT examples, symmetric-heap ring / push all-gather, the two self-tests of ``lk.shmem``
* ``gemm_sm100``            tcgen05 GEMM ladder (``LEVELS`` 1..9: single stage -> rings -> CTA pairs -> persistent, two TMEM accumulators)
* ``ag_gemm`` / ``gemm_rs`` / ``gemm_ar``   the three fused tensor-parallel ops as single kernels (comm CTAs, epilogue reductions, NVLS consumers)
* ``allreduce_nvls`` / ``collectives_nvls`` all-reduce, reduce-scatter, all-gather through the multicast alias (``multimem``)
* ``allreduce_push`` / ``allreduce_tree`` / ``reduce_scatter_ring`` / ``allgather_ll``   peer-to-peer algorithms: one / two shot, double binary tree,
                            ring, flag-in-data atoms
* ``all_to_all`` / ``ep_a2a``   low-latency variable all-to-all on the OpenSHMEM-style API; expert-parallel dispatch / combine
* ``flash_mma`` / ``flash_decode`` / ``linear_mma``   prefill attention and decode linear on mma.sync, split-KV decode + KV-sharded decode
* ``gdn_chunk``             chunked gated-delta-rule forward

Guide: docs/lk.md; generated intrinsic reference: docs/lk_intrinsics.md; micro-benchmarks: ``triton_dist.lk.bench``.
"""
