"""Low-latency all-gather with the flag inside the data (NCCL-LL style), in the DSL.

Reference: kernels/nvidia/low_latency_allgather.py ``_pack_ll_block`` / ``_recv_ll_block`` / ``_forward_push_2d_ll_kernel`` (:531-567,
:700): every 4 bytes of payload travel as an 8-byte atom ``{data, flag}`` written with ONE 64-bit store, so the receiver needs no
separate flag, no fence and no barrier -- it spins on the atom itself until the flag half equals the call number.  Twice the bytes on
the wire for the lowest possible latency: the protocol of choice for the (O, LSE) exchange of distributed flash-decode and other
microsecond-scale gathers.  (The product's CUDA version is mode ``push_2d_ll`` of ``ops.comm.fast_allgather``.)

* send: thread i packs word i of the local shard with the call number and stores the atom into slot ``me`` of every peer's buffer
  (``symm_at``), peers visited in rotated order;
* receive: thread i spins on atom i of every source slot (relaxed 64-bit loads: the store is single-copy atomic, so data and flag arrive
  together), writes the data word to ``out``;
* the buffer is double-buffered by call parity: a sender may be one call ahead of a receiver that is still spinning on the previous
  call's atoms, never two (its own receive loop of call k needs the receiver's call-k atoms).  Flags are call numbers: no reset.

``tests/dist_worker.py`` case ``lk_ag_ll`` runs it across processes in the interpreter against ``torch.distributed.all_gather``.
"""
from triton_dist import lk
from triton_dist.lk import language_extra as le
from triton_dist.lk import ll

THREADS = 256


@lk.kernel(block=THREADS)
def allgather_ll(ctx: ll.SymmCtx, shard: ll.ptr[ll.u32], out: ll.ptr[ll.u32], buf: ll.ptr[ll.u64], nwords: ll.i32, max_words: ll.i32,
                 phase: ll.u32):
    me = ll.rank(ctx)
    W = ll.num_ranks(ctx)
    gid = ll.blockIdx.x * ll.blockDim.x + ll.threadIdx.x
    nthr = ll.gridDim.x * ll.blockDim.x
    half = ll.i64(ll.i32(phase & 1)) * W * max_words
    tag = ll.u64(phase) << 32
    for q in range(W):
        peer = (me + q) % W
        dst = ll.symm_at(ctx, buf, peer) + (half + ll.i64(me) * max_words)
        for i in range(gid, nwords, nthr):
            le.st(dst + i, tag | ll.u64(shard[i]), scope="sys", semantic="relaxed")     # data and flag in ONE 8-byte store
    for src in range(W):
        slot = buf + (half + ll.i64(src) * max_words)
        for i2 in range(gid, nwords, nthr):
            atom = le.ld(slot + i2, scope="sys", semantic="relaxed")
            while (atom >> 32) != ll.u64(phase):
                atom = le.ld(slot + i2, scope="sys", semantic="relaxed")
            out[ll.i64(src) * nwords + i2] = ll.u32(atom & 0xFFFFFFFF)


class LkLLAllGather:
    """Symmetric atom buffer [2, W, max_words] (8 bytes per payload word) for shards of up to ``max_bytes`` (multiple of 4)."""

    def __init__(self, max_bytes: int, grid: int = 0):
        import torch
        import triton_dist.utils as U
        self.W, self.rank = U.world_size(), U.rank()
        self.max_words = (max_bytes + 3) // 4
        self.gpu = U.current_device().type == "cuda"
        self.grid = grid or (min(16, max(1, self.max_words // (4 * THREADS))) if self.gpu else 1)
        self.buf = U.nvshmem_create_tensor((2 * self.W * self.max_words,), torch.int64)
        self.buf.zero_()
        self.phase = 0
        U.barrier_all_on_stream()

    def __call__(self, shard, out=None):
        """shard: contiguous tensor whose byte size is a multiple of 4 -> [W, *shard.shape]."""
        import torch
        nbytes = shard.numel() * shard.element_size()
        assert shard.is_contiguous() and nbytes % 4 == 0 and nbytes // 4 <= self.max_words
        out = torch.empty((self.W,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device) if out is None else out
        self.phase += 1
        args = (lk.symm_ctx(), shard.view(torch.uint8).view(-1).view(torch.int32), out.view(torch.uint8).view(-1).view(torch.int32), self.buf,
                nbytes // 4, self.max_words, self.phase)
        if self.gpu:
            allgather_ll[self.grid](*args)
        else:
            allgather_ll.interpret(self.grid, *args, block=32)            # a warp is enough at interpreter sizes (every thread spins)
        return out

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.buf)
