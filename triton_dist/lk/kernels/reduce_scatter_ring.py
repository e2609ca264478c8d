"""Ring reduce-scatter over the symmetric heap, in the DSL:  out[chunk] = sum over ranks of x_r[me * chunk .. (me + 1) * chunk).

Reference: kernels/nvidia/reduce_scatter.py ``reduce_scatter_ring_push_1d_intra_node_kernel`` (:285-339, the SM ring for boxes without a
switch that reduces: per-segment flags, W - 1 hops).  The product's reduce-scatter pulls through the NVSwitch (``multimem.ld_reduce``,
csrc/comm_kernels.cu) -- this kernel is the peer-to-peer algorithm for topologies where that is not available, and the all-in-Python
example of a multi-step ring protocol:

* step s (0 .. W - 2): rank r forwards chunk ``(r - s - 1) mod W`` to rank r + 1 -- its own contribution, plus (from step 1 on) the
  partial it received in the previous step, which is the same chunk one hop earlier.  After W - 1 steps the chunk that arrives is the
  rank's own, complete except for its own contribution, which is added while writing ``out``.
* every CTA owns a slice of the chunk and runs the whole ring for it on its own flags (``flags[step * grid + cta]``): no grid barrier.
  Flags carry the call number (no reset); receive slots are double-buffered by call parity -- a rank can be at most one call ahead of
  its successor, because finishing a call needs the successor's first send of that call, which the successor issues only after it has
  finished reading the previous call's slots.
* data first, then ``__syncthreads``, then one thread's release store of the flag on the peer; the reader acquires the flag before it
  touches the slot.

fp32 payload (sums are exact to rounding in any order only for fp32 accumulation; a bf16 payload would round at every hop).
``tests/dist_worker.py`` case ``lk_rs_ring`` runs it across processes in the interpreter against ``torch.distributed.reduce_scatter``.
"""
from triton_dist import lk
from triton_dist.lk import ll

THREADS = 256


@lk.kernel(block=THREADS)
def reduce_scatter_ring(ctx: ll.SymmCtx, x: ll.ptr[ll.f32], recv: ll.ptr[ll.f32], flags: ll.ptr[ll.u32], out: ll.ptr[ll.f32],
                        chunk: ll.i32, phase: ll.u32):
    me = ll.rank(ctx)
    W = ll.num_ranks(ctx)
    G = ll.gridDim.x
    cta = ll.blockIdx.x
    tid = ll.threadIdx.x
    per = (chunk + G - 1) // G
    lo = cta * per
    hi = min(chunk, lo + per)
    nxt = (me + 1) % W
    slot0 = ll.i64(ll.i32(phase & 1) * (W - 1)) * chunk                 # this call's W - 1 receive slots
    for step in range(W - 1):
        c = (me - step - 1 + 2 * W) % W
        if step > 0:
            if tid < 32:
                ll.wait(flags + ((step - 1) * G + cta), 1, phase, True)    # acquire: the previous hop's partial is in my slot
            ll.syncthreads()
        dst = ll.symm_at(ctx, recv, nxt) + (slot0 + ll.i64(step) * chunk)
        for i in range(lo + tid, hi, THREADS):
            v = x[ll.i64(c) * chunk + i]
            if step > 0:
                v += recv[slot0 + ll.i64(step - 1) * chunk + i]
            dst[i] = v
        ll.syncthreads()                                                  # the slice is stored ...
        if tid == 0:
            ll.notify(ctx, flags + (step * G + cta), nxt, phase)          # ... then the release store of the flag on the successor
    if W > 1:
        if tid < 32:
            ll.wait(flags + ((W - 2) * G + cta), 1, phase, True)
        ll.syncthreads()
    for i2 in range(lo + tid, hi, THREADS):
        acc = x[ll.i64(me) * chunk + i2]
        if W > 1:
            acc += recv[slot0 + ll.i64(W - 2) * chunk + i2]
        out[i2] = acc


class LkRingReduceScatter:
    """Symmetric receive slots [2, W - 1, chunk] + flags [(W - 1) * grid] for chunks of up to ``max_chunk`` fp32 elements."""

    def __init__(self, max_chunk: int, grid: int = 0):
        import torch
        import triton_dist.utils as U
        self.W, self.rank = U.world_size(), U.rank()
        self.max_chunk = max_chunk
        self.gpu = U.current_device().type == "cuda"
        self.grid = grid or (min(32, max(1, max_chunk // 4096)) if self.gpu else 2)
        self.recv = U.nvshmem_create_tensor((2 * max(self.W - 1, 1) * max_chunk,), torch.float32)
        self.flags = U.nvshmem_create_tensor((max(self.W - 1, 1) * self.grid,), torch.int32)
        self.flags.zero_()
        self.phase = 0
        U.barrier_all_on_stream()

    def __call__(self, x, out=None):
        """x: [W * chunk] (or [W, chunk]) fp32 -> this rank's reduced chunk [chunk]."""
        import torch
        chunk = x.numel() // self.W
        assert x.dtype == torch.float32 and x.numel() == chunk * self.W and chunk <= self.max_chunk
        out = torch.empty(chunk, dtype=torch.float32, device=x.device) if out is None else out
        self.phase += 1
        # the kernel indexes the receive slots with the ACTUAL chunk length: slot (parity, step) starts at ((parity * (W - 1)) + step) * chunk
        args = (lk.symm_ctx(), x.contiguous().view(-1), self.recv, self.flags, out, chunk, self.phase)
        if self.gpu:
            reduce_scatter_ring[self.grid](*args)
        else:
            reduce_scatter_ring.interpret(self.grid, *args)
        return out

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.flags)
        U.nvshmem_free_tensor_sync(self.recv)
