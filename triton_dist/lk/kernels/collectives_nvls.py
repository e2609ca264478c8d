"""Reduce-scatter and all-gather through the NVSwitch (NVLS), in the DSL -- the two halves of the two-shot all-reduce as collectives of
their own.

Reference: the product's ``reduce_scatter`` (csrc/comm_kernels.cu method 5: every rank ``multimem.ld_reduce``s ITS rows) and the NVLS
all-gather (``allgather_mc_kernel``: one ``multimem.st`` stream per rank instead of W - 1 unicast copies; reference
low_latency_allgather.py ``_forward_push_2d_ll_multimem_kernel`` without the LL packing):

* ``reduce_scatter_nvls``: stage ``x`` (W chunks) in the symmetric buffer, barrier, reduce chunk ``me`` through the multicast alias into
  ``out``, exit barrier;
* ``allgather_nvls``: ``multimem.st`` the local shard into slot ``me`` of EVERY rank's buffer (the switch replicates: egress is 1 x the
  shard), barrier, copy the W slots out, exit barrier.

bf16 / fp32 payloads in 16-byte vectors; one flag-flip team barrier state per CTA (``lk.shmem``).  Both run in the interpreter with the
multicast model (``tests/dist_worker.py`` case ``lk_nvls_collectives``).
"""
from triton_dist import lk
from triton_dist.lk import ll, shmem

THREADS = 256


def make_nvls_collectives(dtype=ll.bf16):
    assert dtype in (ll.bf16, ll.f32)
    VEC = 8 if dtype is ll.bf16 else 4

    def ld_reduce(mc):
        if VEC == 8:
            return ll.multimem_ld_reduce_bf16x8(mc)
        return ll.multimem_ld_reduce_f32x4(mc)

    @lk.kernel(block=THREADS)
    def reduce_scatter_nvls(ctx: ll.SymmCtx, x: ll.ptr[dtype], out: ll.ptr[dtype], buf: ll.ptr[dtype], slots: ll.ptr[ll.u32],
                            epoch: ll.ptr[ll.u32], chunk_vec: ll.i32):
        """x: W chunks of ``chunk_vec`` 16-byte vectors -> out: the sum over ranks of chunk ``me``."""
        G = ll.gridDim.x
        cta = ll.blockIdx.x
        tid = ll.threadIdx.x
        W = ll.num_ranks(ctx)
        me = ll.rank(ctx)
        s = shmem.make_sync(slots + cta * 2 * W, epoch + cta)
        for v in range(cta * THREADS + tid, chunk_vec * W, G * THREADS):
            ll.st_v4(buf + ll.i64(v) * VEC, ll.ld_v4(x + ll.i64(v) * VEC))
        shmem.barrier_all_block(ctx, s)
        mc = ll.symm_mc(ctx, buf) + ll.i64(me) * chunk_vec * VEC
        for v2 in range(cta * THREADS + tid, chunk_vec, G * THREADS):
            ll.st_v4(out + ll.i64(v2) * VEC, ld_reduce(mc + ll.i64(v2) * VEC))
        shmem.barrier_all_block(ctx, s)

    @lk.kernel(block=THREADS)
    def allgather_nvls(ctx: ll.SymmCtx, shard: ll.ptr[dtype], out: ll.ptr[dtype], buf: ll.ptr[dtype], slots: ll.ptr[ll.u32],
                       epoch: ll.ptr[ll.u32], shard_vec: ll.i32):
        """shard: ``shard_vec`` 16-byte vectors -> out: [W, shard]."""
        G = ll.gridDim.x
        cta = ll.blockIdx.x
        tid = ll.threadIdx.x
        W = ll.num_ranks(ctx)
        me = ll.rank(ctx)
        s = shmem.make_sync(slots + cta * 2 * W, epoch + cta)
        mc = ll.symm_mc(ctx, buf) + ll.i64(me) * shard_vec * VEC
        for v in range(cta * THREADS + tid, shard_vec, G * THREADS):
            ll.multimem_st_v4(mc + ll.i64(v) * VEC, ll.ld_v4(shard + ll.i64(v) * VEC))      # ONE store, every rank's slot `me`
        shmem.barrier_all_block(ctx, s)
        for v2 in range(cta * THREADS + tid, shard_vec * W, G * THREADS):
            ll.st_v4(out + ll.i64(v2) * VEC, ll.ld_v4(buf + ll.i64(v2) * VEC))
        shmem.barrier_all_block(ctx, s)

    tag = "bf16" if VEC == 8 else "f32"
    reduce_scatter_nvls.name, allgather_nvls.name = f"lk_reduce_scatter_nvls_{tag}", f"lk_allgather_nvls_{tag}"
    return reduce_scatter_nvls, allgather_nvls


_KERNELS = {}


class LkNvlsCollectives:
    """``reduce_scatter(x)`` / ``all_gather(shard)`` on one symmetric staging buffer of ``max_bytes`` (the full gathered / unreduced size)."""

    def __init__(self, max_bytes: int, grid: int = 1):
        import torch
        import triton_dist.utils as U
        assert max_bytes % 16 == 0
        self.W, self.grid, self.max_bytes = U.world_size(), grid, max_bytes
        dev = U.current_device()
        self.gpu = dev.type == "cuda"
        if self.gpu and not U.is_nvshmem_multimem_supported():
            raise RuntimeError("the NVLS collectives need the multicast mapping of the symmetric heap")
        self.buf = U.nvshmem_create_tensor((max_bytes,), torch.uint8)
        self.slots = U.nvshmem_create_tensor((grid * 2 * self.W,), torch.int32)
        self.slots.zero_()
        self.epoch = torch.zeros(grid, dtype=torch.int32, device=dev)
        U.barrier_all_on_stream()

    def _kernels(self, dt):
        import torch
        if dt not in _KERNELS:
            _KERNELS[dt] = make_nvls_collectives(ll.bf16 if dt == torch.bfloat16 else ll.f32)
        return _KERNELS[dt]

    def _run(self, k, *args):
        if self.gpu:
            k[self.grid](*args)
        else:
            k.interpret(self.grid, *args)

    def reduce_scatter(self, x, out=None):
        """x: [W * chunk] (bf16 / fp32, chunk bytes a multiple of 16) -> [chunk] = sum over ranks of chunk ``me``."""
        import torch
        nbytes = x.numel() * x.element_size()
        assert x.is_contiguous() and nbytes <= self.max_bytes and nbytes % (16 * self.W) == 0
        out = torch.empty(x.numel() // self.W, dtype=x.dtype, device=x.device) if out is None else out
        self._run(self._kernels(x.dtype)[0], lk.symm_ctx(), x.view(-1), out, self.buf.view(x.dtype), self.slots, self.epoch, nbytes // 16 // self.W)
        return out

    def all_gather(self, shard, out=None):
        """shard: contiguous, bytes a multiple of 16 -> [W, *shard.shape]."""
        import torch
        nbytes = shard.numel() * shard.element_size()
        assert shard.is_contiguous() and nbytes % 16 == 0 and nbytes * self.W <= self.max_bytes
        out = torch.empty((self.W,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device) if out is None else out
        self._run(self._kernels(shard.dtype)[1], lk.symm_ctx(), shard.view(-1), out.view(-1), self.buf.view(shard.dtype), self.slots, self.epoch, nbytes // 16)
        return out

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.slots)
        U.nvshmem_free_tensor_sync(self.buf)
