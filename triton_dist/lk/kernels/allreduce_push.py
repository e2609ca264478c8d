"""One-shot and two-shot PUSH all-reduce over the symmetric heap, in the DSL (fp32).

Reference: kernels/nvidia/allreduce.py ``allreduce_one_shot_push_intra_node_kernel`` (:334-384: put to every peer, then sum the W copies
locally) and ``allreduce_two_shot_push_intra_node_kernel`` (:448-525: reduce-scatter by put + local reduce + all-gather by put).  The
product's kernels pull instead (peers read a rank's staged input over NVLink or through the multicast mapping, csrc/comm_kernels.cu); these
are the push formulations as two short Python kernels:

* one shot -- W x the bytes on the wire, ONE flag round: every CTA stores its slice of ``x`` into slot ``me`` of every peer's buffer,
  release-flags ``flags[cta][me]`` there with the call number, waits for its own W flags and sums the W slots.  Best for small messages.
* two shot -- 2 (W - 1) / W x the bytes, two flag rounds: slice s of ``x`` goes to rank s (slot ``me``), rank s reduces its slice and
  pushes the result into every rank's ``out`` (a symmetric tensor) -- a reduce-scatter and an all-gather back to back.

Buffers are double-buffered by call parity (a rank can be at most one call ahead of a peer: finishing a call needs every peer's push of
that call, which a peer issues only after it has finished reading the previous call's slots); flags are never reset.
``tests/dist_worker.py`` case ``lk_ar_push`` runs both across processes in the interpreter against ``torch.distributed.all_reduce``.
"""
from triton_dist import lk
from triton_dist.lk import ll

THREADS = 256


@lk.kernel(block=THREADS)
def allreduce_one_shot_push(ctx: ll.SymmCtx, x: ll.ptr[ll.f32], out: ll.ptr[ll.f32], buf: ll.ptr[ll.f32], flags: ll.ptr[ll.u32],
                            n: ll.i32, max_n: ll.i32, phase: ll.u32):
    me = ll.rank(ctx)
    W = ll.num_ranks(ctx)
    G = ll.gridDim.x
    cta = ll.blockIdx.x
    tid = ll.threadIdx.x
    per = (n + G - 1) // G
    lo = cta * per
    hi = min(n, lo + per)
    half = ll.i64(ll.i32(phase & 1)) * W * max_n                           # this call's W slots
    for q in range(W):
        peer = (me + q) % W                                              # every rank starts with a different destination
        dst = ll.symm_at(ctx, buf, peer) + (half + ll.i64(me) * max_n)
        for i in range(lo + tid, hi, THREADS):
            dst[i] = x[i]
        ll.syncthreads()
        if tid == 0:
            ll.notify(ctx, flags + (cta * W + me), peer, phase)
    if tid < 32:
        ll.wait(flags + cta * W, W, phase, True)                          # all W sources have delivered this CTA's slice
    ll.syncthreads()
    for i2 in range(lo + tid, hi, THREADS):
        acc: ll.f32 = 0.0
        for r in range(W):
            acc += buf[half + ll.i64(r) * max_n + i2]
        out[i2] = acc


@lk.kernel(block=THREADS)
def allreduce_two_shot_push(ctx: ll.SymmCtx, x: ll.ptr[ll.f32], out: ll.ptr[ll.f32], buf: ll.ptr[ll.f32], flags: ll.ptr[ll.u32],
                            seg: ll.i32, max_seg: ll.i32, phase: ll.u32):
    """``x`` / ``out``: W segments of ``seg`` elements; ``out`` is SYMMETRIC (peers write into it)."""
    me = ll.rank(ctx)
    W = ll.num_ranks(ctx)
    G = ll.gridDim.x
    cta = ll.blockIdx.x
    tid = ll.threadIdx.x
    per = (seg + G - 1) // G
    lo = cta * per
    hi = min(seg, lo + per)
    half = ll.i64(ll.i32(phase & 1)) * W * max_seg
    f_rs = flags + cta * W                                               # reduce-scatter arrivals, one per source
    f_ag = flags + (G + cta) * W                                         # all-gather arrivals, one per owner
    # ---- reduce-scatter by push: segment s of x -> slot `me` on rank s ----
    for q in range(W):
        owner = (me + q) % W
        dst = ll.symm_at(ctx, buf, owner) + (half + ll.i64(me) * max_seg)
        for i in range(lo + tid, hi, THREADS):
            dst[i] = x[ll.i64(owner) * seg + i]
        ll.syncthreads()
        if tid == 0:
            ll.notify(ctx, f_rs + me, owner, phase)
    if tid < 32:
        ll.wait(f_rs, W, phase, True)
    ll.syncthreads()
    # ---- reduce my segment, all-gather by push: the result -> segment `me` of everyone's out ----
    for q2 in range(W):
        peer = (me + q2) % W
        o = ll.symm_at(ctx, out, peer) + ll.i64(me) * seg
        for i2 in range(lo + tid, hi, THREADS):
            acc: ll.f32 = 0.0
            for r in range(W):
                acc += buf[half + ll.i64(r) * max_seg + i2]
            o[i2] = acc
        ll.syncthreads()
        if tid == 0:
            ll.notify(ctx, f_ag + me, peer, phase)
    if tid < 32:
        ll.wait(f_ag, W, phase, True)                                     # every owner's segment has landed in my out
    ll.syncthreads()


class LkPushAllReduce:
    """``method``: ``"one_shot"`` or ``"two_shot"``; messages of up to ``max_elems`` fp32 elements (two shot: a multiple of the world
    size).  ``__call__`` returns the sum over all ranks; the two-shot result is a view of a symmetric tensor owned by this object, valid until the
    next call (peers write the next result into it only after this rank has entered its next call)."""

    def __init__(self, max_elems: int, method: str = "one_shot", grid: int = 0):
        import torch
        import triton_dist.utils as U
        assert method in ("one_shot", "two_shot")
        self.method, self.W, self.rank = method, U.world_size(), U.rank()
        self.gpu = U.current_device().type == "cuda"
        self.unit = max_elems if method == "one_shot" else (max_elems + self.W - 1) // self.W        # slot length
        self.grid = grid or (min(32, max(1, self.unit // 4096)) if self.gpu else 2)
        self.buf = U.nvshmem_create_tensor((2 * self.W * self.unit,), torch.float32)
        self.flags = U.nvshmem_create_tensor((2 * self.grid * self.W,), torch.int32)
        self.out = U.nvshmem_create_tensor((self.W * self.unit,), torch.float32) if method == "two_shot" else None
        self.flags.zero_()
        self.phase = 0
        U.barrier_all_on_stream()

    def __call__(self, x, out=None):
        import torch
        n = x.numel()
        assert x.dtype == torch.float32
        self.phase += 1
        if self.method == "one_shot":
            assert n <= self.unit
            out = torch.empty_like(x) if out is None else out
            args = (lk.symm_ctx(), x.contiguous().view(-1), out.view(-1), self.buf, self.flags, n, self.unit, self.phase)
            k = allreduce_one_shot_push
        else:
            assert n % self.W == 0 and n // self.W <= self.unit
            args = (lk.symm_ctx(), x.contiguous().view(-1), self.out, self.buf, self.flags, n // self.W, self.unit, self.phase)
            k = allreduce_two_shot_push
        if self.gpu:
            k[self.grid](*args)
        else:
            k.interpret(self.grid, *args)
        if self.method == "two_shot":
            res = self.out[:n].view(x.shape)
            return res if out is None else out.copy_(res)
        return out

    def finalize(self):
        import triton_dist.utils as U
        for t in (self.out, self.flags, self.buf):
            if t is not None:
                U.nvshmem_free_tensor_sync(t)
