"""All-reduce through the NVSwitch (NVLS multicast), in the DSL: the one-shot and two-shot ``multimem`` algorithms.

Reference: kernels/nvidia/allreduce.py ``allreduce_one_shot_multimem_intra_node_kernel`` (:603-657) and
``allreduce_two_shot_multimem_intra_node_kernel`` (:661-683).  These are the algorithms behind the product's default all-reduce
(csrc/comm_kernels.cu methods 2 / 3, hardware-validated); here they are two short Python kernels on the same PTX wrappers
(``multimem.ld_reduce`` / ``multimem.st`` through ``ll.symm_mc``), which also run in the CPU interpreter -- a multimem load there reads the
word from every rank's copy on the emulation heap and adds, exactly what the switch does:

* one shot: stage ``x`` in the symmetric buffer, cross-rank barrier, then EVERY rank reduces the WHOLE message with 16-byte
  ``multimem.ld_reduce`` loads (the switch adds the W copies in flight) and stores to its own ``out``; exit barrier before the buffer may
  be overwritten.  W x the reduction traffic through the switch, one pass -- best for small messages.
* two shot: after the same staging + barrier every rank reduces only ITS 1 / W slice with ``multimem.ld_reduce`` and broadcasts the result
  into everybody's buffer with ``multimem.st``; second barrier; copy out.  Each byte crosses the switch twice, independent of W.

bf16 or fp32 payloads (``multimem.ld_reduce...bf16x2`` accumulates in fp32 inside the switch).  The barriers are the flag-flip team barrier
of ``lk.shmem`` (one CTA: the kernels are latency-bound; grid > 1 splits the message and every CTA synchronises on its own slots).
``tests/dist_worker.py`` case ``lk_ar_nvls`` checks both against ``torch.distributed.all_reduce`` across processes.
"""
from triton_dist import lk
from triton_dist.lk import ll, shmem

THREADS = 256


def make_allreduce_nvls(dtype=ll.bf16):
    """Kernel pair for one payload type (``ll.bf16`` or ``ll.f32``): the 16-byte vector holds 8 or 4 elements."""
    assert dtype in (ll.bf16, ll.f32)
    VEC = 8 if dtype is ll.bf16 else 4

    def ld_reduce(mc):
        if VEC == 8:
            return ll.multimem_ld_reduce_bf16x8(mc)
        return ll.multimem_ld_reduce_f32x4(mc)

    @lk.kernel(block=THREADS)
    def one_shot(ctx: ll.SymmCtx, x: ll.ptr[dtype], out: ll.ptr[dtype], buf: ll.ptr[dtype], slots: ll.ptr[ll.u32], epoch: ll.ptr[ll.u32],
                 nvec: ll.i32):
        G = ll.gridDim.x
        cta = ll.blockIdx.x
        tid = ll.threadIdx.x
        W = ll.num_ranks(ctx)
        s = shmem.make_sync(slots + cta * 2 * W, epoch + cta)
        per = (nvec + G - 1) // G
        lo = cta * per
        hi = min(nvec, lo + per)
        for v in range(lo + tid, hi, THREADS):
            ll.st_v4(buf + ll.i64(v) * VEC, ll.ld_v4(x + ll.i64(v) * VEC))
        shmem.barrier_all_block(ctx, s)                                   # everybody's input is staged and visible
        mc = ll.symm_mc(ctx, buf)
        for v2 in range(lo + tid, hi, THREADS):
            ll.st_v4(out + ll.i64(v2) * VEC, ld_reduce(mc + ll.i64(v2) * VEC))
        shmem.barrier_all_block(ctx, s)                                   # nobody restages while a peer still reduces

    @lk.kernel(block=THREADS)
    def two_shot(ctx: ll.SymmCtx, x: ll.ptr[dtype], out: ll.ptr[dtype], buf: ll.ptr[dtype], slots: ll.ptr[ll.u32], epoch: ll.ptr[ll.u32],
                 nvec: ll.i32):
        G = ll.gridDim.x
        cta = ll.blockIdx.x
        tid = ll.threadIdx.x
        W = ll.num_ranks(ctx)
        me = ll.rank(ctx)
        s = shmem.make_sync(slots + cta * 2 * W, epoch + cta)
        per = (nvec + G - 1) // G
        lo = cta * per
        hi = min(nvec, lo + per)
        for v in range(lo + tid, hi, THREADS):
            ll.st_v4(buf + ll.i64(v) * VEC, ll.ld_v4(x + ll.i64(v) * VEC))
        shmem.barrier_all_block(ctx, s)
        mc = ll.symm_mc(ctx, buf)
        # my share of this CTA's range: vectors lo + me, lo + me + W, ... (interleaved so that every rank owns some of any range)
        for v2 in range(lo + me + tid * W, hi, THREADS * W):
            r = ld_reduce(mc + ll.i64(v2) * VEC)
            ll.multimem_st_v4(mc + ll.i64(v2) * VEC, r)                   # the reduced vector replaces the input in EVERY rank's buffer
        shmem.barrier_all_block(ctx, s)                                   # all owners have broadcast
        for v3 in range(lo + tid, hi, THREADS):
            ll.st_v4(out + ll.i64(v3) * VEC, ll.ld_v4(buf + ll.i64(v3) * VEC))
        shmem.barrier_all_block(ctx, s)

    tag = "bf16" if VEC == 8 else "f32"
    one_shot.name, two_shot.name = f"lk_allreduce_nvls_one_shot_{tag}", f"lk_allreduce_nvls_two_shot_{tag}"
    return one_shot, two_shot


_KERNELS = {}


class LkNvlsAllReduce:
    """Symmetric staging buffer + per-CTA barrier state for messages of up to ``max_bytes`` (multiple of 16)."""

    def __init__(self, max_bytes: int, method: str = "one_shot", grid: int = 1):
        import torch
        import triton_dist.utils as U
        assert method in ("one_shot", "two_shot") and max_bytes % 16 == 0
        self.method, self.W, self.grid = method, U.world_size(), grid
        dev = U.current_device()
        self.gpu = dev.type == "cuda"
        if self.gpu and not U.is_nvshmem_multimem_supported():
            raise RuntimeError("the NVLS all-reduce kernels need the multicast mapping of the symmetric heap")
        self.buf = U.nvshmem_create_tensor((max_bytes,), torch.uint8)
        self.slots = U.nvshmem_create_tensor((grid * 2 * self.W,), torch.int32)
        self.slots.zero_()
        self.epoch = torch.zeros(grid, dtype=torch.int32, device=dev)
        self.max_bytes = max_bytes
        U.barrier_all_on_stream()

    def __call__(self, x, out=None):
        import torch
        assert x.dtype in (torch.bfloat16, torch.float32) and x.is_contiguous()
        nbytes = x.numel() * x.element_size()
        assert nbytes % 16 == 0 and nbytes <= self.max_bytes
        key = x.dtype
        if key not in _KERNELS:
            _KERNELS[key] = make_allreduce_nvls(ll.bf16 if x.dtype == torch.bfloat16 else ll.f32)
        k = _KERNELS[key][0 if self.method == "one_shot" else 1]
        out = torch.empty_like(x) if out is None else out
        args = (lk.symm_ctx(), x.view(-1), out.view(-1), self.buf.view(x.dtype), self.slots, self.epoch, nbytes // 16)
        if self.gpu:
            k[self.grid](*args)
        else:
            k.interpret(self.grid, *args)
        return out

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.slots)
        U.nvshmem_free_tensor_sync(self.buf)
