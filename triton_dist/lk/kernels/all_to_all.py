"""Low-latency variable-size all-to-all written with the OpenSHMEM-style device API of the DSL (``triton_dist.lk.shmem``).

Reference: kernels/nvidia/low_latency_all_to_all.py ``all_to_all_kernel`` (:33-119; tutorials/04-deepseek-infer-all2all): grid = (world,),
CTA ``pid`` sends this rank's rows for PE ``pid`` with ``putmem_nbi_block`` + the row count with ``putmem_signal_nbi_block``; buffers are
double-buffered by ``call_count % 2`` and the signal value is the call count; the CTA then waits for the signal of source ``pid``.
The same structure, one Python function:

* CTA d: ``shmem.putmem_block`` of rows ``[cum[d], cum[d + 1])`` into slot ``me`` of PE d's receive buffer (parity half), then
  ``shmem.putmem_signal_block`` of the row count with the call number as signal value (data first, signal last, release semantics);
* the same CTA then ``signal_wait_until(sig[d] >= call)``: source d's rows and count for this rank have landed.

``LkAllToAll`` returns views of the receive buffer ``[W, max_rows, H]`` and the received counts ``[W]`` -- valid until the call after next.
``tests/dist_worker.py`` case ``lk_a2a`` checks it against ``torch.distributed.all_to_all_single`` with uneven splits across processes.
"""
from triton_dist import lk
from triton_dist.lk import ll, shmem

THREADS = 256


@lk.kernel(block=THREADS)
def all_to_all_ll(ctx: ll.SymmCtx, send: ll.ptr[ll.bf16], cum: ll.ptr[ll.i32], recv: ll.ptr[ll.bf16], counts: ll.ptr[ll.i32],
                  sig: ll.ptr[ll.u64], my_counts: ll.ptr[ll.i32], H: ll.i32, max_rows: ll.i32, phase: ll.u64):
    """send: [total_rows, H]; cum: int32 [W + 1] row offsets per destination; recv: symmetric [2, W, max_rows, H]; counts: symmetric
    int32 [2, W]; sig: symmetric uint64 [W]; my_counts: local int32 [W] scratch (row count per destination, source of the count put)."""
    W = shmem.n_pes(ctx)
    me = shmem.my_pe(ctx)
    d = ll.blockIdx.x
    par = ll.i32(phase & 1)
    r0 = cum[d]
    n = cum[d + 1] - r0
    if ll.threadIdx.x == 0:
        my_counts[d] = n
    ll.syncthreads()
    slot = recv + (ll.i64(par * W + me) * max_rows) * H
    shmem.putmem_block(ctx, slot, send + ll.i64(r0) * H, ll.i64(n) * H * 2, d)
    shmem.putmem_signal_block(ctx, counts + (par * W + me), my_counts + d, 4, sig + me, phase, shmem.SIGNAL_SET, d)
    if ll.threadIdx.x == 0:
        shmem.signal_wait_until(sig + d, shmem.CMP_GE, phase)        # source d's rows + count for me are visible after this
    ll.syncthreads()


class LkAllToAll:
    def __init__(self, max_rows: int, hidden: int):
        import torch
        import triton_dist.utils as U
        self.W, self.rank, self.max_rows, self.H = U.world_size(), U.rank(), max_rows, hidden
        self.gpu = U.current_device().type == "cuda"
        self.recv = U.nvshmem_create_tensor((2, self.W, max_rows, hidden), torch.bfloat16)
        self.counts = U.nvshmem_create_tensor((2, max(self.W, 2)), torch.int32)
        self.sig = U.nvshmem_create_tensor((max(self.W, 2),), torch.int64)
        self.sig.zero_()
        self.my_counts = torch.zeros(self.W, dtype=torch.int32, device=U.current_device())
        self.phase = 0
        U.barrier_all_on_stream()

    def __call__(self, send, splits):
        """send: [sum(splits), H] bf16, rows grouped by destination; splits: int32 [W] -> (recv [W, max_rows, H], recv_counts [W])."""
        import torch
        assert send.dtype == torch.bfloat16 and send.shape[1] == self.H and splits.numel() == self.W
        cum = torch.zeros(self.W + 1, dtype=torch.int32, device=send.device)
        cum[1:] = torch.cumsum(splits.to(torch.int32), 0)
        self.phase += 1
        par = self.phase & 1
        args = (lk.symm_ctx(), send.contiguous(), cum, self.recv, self.counts, self.sig, self.my_counts, self.H, self.max_rows, self.phase)
        if self.gpu:
            all_to_all_ll[self.W](*args)
        else:
            all_to_all_ll.interpret(self.W, *args)
        return self.recv[par], self.counts[par, : self.W]

    def finalize(self):
        import triton_dist.utils as U
        for t in (self.sig, self.counts, self.recv):
            U.nvshmem_free_tensor_sync(t)
