"""Expert-parallel dispatch / combine written in the DSL (one NVSwitch domain).

Reference: little_kernel/design/flashcomm_{compute,dispatch,combine,postprocess,ep_kernels}.py -- the FlashComm EP all-to-all rebuilt with
the reference's DSL (host manager ``FlashCommEPKernels``).  The production EP kernels of this framework are CUDA (csrc/ep_kernels.cu,
csrc/ep_normal_kernels.cu, the Mega-EP modes of the GEMM); this module is the same capability as two short Python kernels, to show that
the DSL plus its symmetric-heap vocabulary is enough for a token-routing protocol, and to have one that runs in the CPU interpreter
across processes:

* ``dispatch``: a warp per (token, k) pair.  The destination rank is ``expert // experts_per_rank``; the row's slot inside the region
  that destination reserves for THIS source comes from a warp-aggregated atomic on a LOCAL counter (every source owns ``cap`` rows on
  every destination, so no remote atomics and no inter-source ordering); the row (16-byte vectors over NVLink) and its (token, local
  expert) tag are stored to the peer, the slot is remembered for the way back.  After a grid barrier, W threads of block 0 publish the
  per-destination row counts and raise ``flags[me]`` on each peer with release semantics (phase-numbered: no reset), then the block waits
  for all W sources.  The kernel's end therefore means: every row addressed to this rank has landed.
* ``combine``: a warp per token pulls the k processed rows back from the peers' buffers (same slots), scales by the routing weights,
  accumulates in fp32 and stores the token.  The caller puts one cross-rank barrier between "experts wrote their outputs" and combine.

``LkEpAllToAll`` owns the symmetric buffers; ``tests/dist_worker.py`` case ``lk_ep`` checks dispatch -> per-expert scaling -> combine
against the dense formula on both backends.
"""
from triton_dist import lk
from triton_dist.lk import language_extra as le
from triton_dist.lk import ll, stdlib

WARPS = 4
THREADS = WARPS * 32


@lk.kernel(block=THREADS)
def ep_dispatch(ctx: ll.SymmCtx, x: ll.ptr[ll.bf16], topk_ids: ll.ptr[ll.i32], recv_x: ll.ptr[ll.bf16], recv_meta: ll.ptr[ll.i32],
                send_count: ll.ptr[ll.u32], send_slot: ll.ptr[ll.i32], recv_count: ll.ptr[ll.u32], flags: ll.ptr[ll.u32],
                bar: ll.ptr[ll.u32], T: ll.i32, H: ll.i32, topk: ll.i32, epr: ll.i32, cap: ll.i32, phase: ll.u32):
    me = ll.rank(ctx)
    W = ll.num_ranks(ctx)
    tid = ll.threadIdx.x
    lane = tid % 32
    warp_global = ll.blockIdx.x * WARPS + tid // 32
    n_warps = ll.gridDim.x * WARPS
    hvec = H // 8
    for pair in range(warp_global, T * topk, n_warps):
        tok = pair // topk
        e = topk_ids[pair]                                       # the same value in every lane: no divergence around the warp atomics
        if e >= 0:
            dst = e // epr
            slot = ll.i32(le.atomic_add_per_warp(send_count + dst, 1))
            keep = slot < cap                                    # rows beyond the reserved region are dropped (slot -1 on the way back)
            if keep:
                row = ll.symm_at(ctx, recv_x, dst) + (ll.i64(me) * cap + slot) * H
                for v in range(lane, hvec, 32):
                    ll.st_v4(row + v * 8, ll.ld_v4(x + (ll.i64(tok) * H + v * 8)))
            if lane == 0:
                if keep:
                    meta = ll.symm_at(ctx, recv_meta, dst) + (ll.i64(me) * cap + slot) * 2
                    meta[0] = tok
                    meta[1] = e % epr
                    send_slot[pair] = slot
                else:
                    send_slot[pair] = -1
        else:
            if lane == 0:
                send_slot[pair] = -1
    stdlib.grid_barrier(bar, ll.gridDim.x * phase)               # every row of this rank has been issued
    if ll.blockIdx.x == 0:
        if tid < W:
            c = min(send_count[tid], ll.u32(cap))
            cnt = ll.symm_at(ctx, recv_count, tid)
            cnt[me] = c
            send_count[tid] = 0                                  # ready for the next call
            ll.notify(ctx, flags + me, tid, phase)               # release: rows, tags and the count are visible to whoever acquires
        if tid < 32:
            ll.wait(flags, W, phase)
        ll.syncthreads()


@lk.kernel(block=THREADS)
def ep_combine(ctx: ll.SymmCtx, y_buf: ll.ptr[ll.bf16], topk_ids: ll.ptr[ll.i32], topk_w: ll.ptr[ll.f32], send_slot: ll.ptr[ll.i32],
               out: ll.ptr[ll.bf16], T: ll.i32, H: ll.i32, topk: ll.i32, epr: ll.i32, cap: ll.i32):
    me = ll.rank(ctx)
    tid = ll.threadIdx.x
    lane = tid % 32
    warp_global = ll.blockIdx.x * WARPS + tid // 32
    n_warps = ll.gridDim.x * WARPS
    hvec = H // 8
    acc = ll.local([8], ll.f32)
    for tok in range(warp_global, T, n_warps):
        for v in range(lane, hvec, 32):
            for j in ll.static_range(8):
                acc[j] = 0.0
            for k in range(topk):
                pair = tok * topk + k
                e = topk_ids[pair]
                slot = send_slot[pair]
                if e >= 0 and slot >= 0:
                    row = ll.symm_at(ctx, y_buf, e // epr) + (ll.i64(me) * cap + slot) * H
                    q = ll.ld_v4(row + v * 8)
                    w = topk_w[pair]
                    acc[0] += w * ll.bf16_lo(q.x)
                    acc[1] += w * ll.bf16_hi(q.x)
                    acc[2] += w * ll.bf16_lo(q.y)
                    acc[3] += w * ll.bf16_hi(q.y)
                    acc[4] += w * ll.bf16_lo(q.z)
                    acc[5] += w * ll.bf16_hi(q.z)
                    acc[6] += w * ll.bf16_lo(q.w)
                    acc[7] += w * ll.bf16_hi(q.w)
            ll.st_v4(out + (ll.i64(tok) * H + v * 8),
                     ll.make_uint4(ll.pack_bf16x2(acc[0], acc[1]), ll.pack_bf16x2(acc[2], acc[3]), ll.pack_bf16x2(acc[4], acc[5]),
                                   ll.pack_bf16x2(acc[6], acc[7])))


class LkEpAllToAll:
    """Symmetric buffers + call counter of the two kernels.  ``cap``: rows every source may send to one destination per call."""

    def __init__(self, max_tokens: int, hidden: int, topk: int, num_experts: int, cap: int = 0, grid: int = 0):
        import torch
        import triton_dist.utils as U
        assert hidden % 8 == 0
        self.W, self.rank = U.world_size(), U.rank()
        assert num_experts % self.W == 0
        self.T, self.H, self.topk, self.E, self.epr = max_tokens, hidden, topk, num_experts, num_experts // self.W
        self.cap = cap or max_tokens * topk
        dev = U.current_device()
        self.recv_x = U.nvshmem_create_tensor((self.W, self.cap, hidden), torch.bfloat16)
        self.y_buf = U.nvshmem_create_tensor((self.W, self.cap, hidden), torch.bfloat16)
        self.recv_meta = U.nvshmem_create_tensor((self.W, self.cap, 2), torch.int32)
        self.recv_count = U.nvshmem_create_tensor((max(self.W, 4),), torch.int32)
        self.flags = U.nvshmem_create_tensor((max(self.W, 4),), torch.int32)
        self.send_count = torch.zeros(max(self.W, 4), dtype=torch.int32, device=dev)
        self.send_slot = torch.zeros(max_tokens * topk, dtype=torch.int32, device=dev)
        self.bar = torch.zeros(1, dtype=torch.int32, device=dev)
        for t in (self.recv_count, self.flags):
            t.zero_()
        self.phase = 0
        self.gpu = dev.type == "cuda"
        self.grid = grid or (min(64, max(1, (max_tokens * topk + WARPS - 1) // WARPS)) if self.gpu else 2)
        U.barrier_all_on_stream()

    def _run(self, k, *args):
        if self.gpu:
            k[self.grid](*args)
        else:
            k.interpret(self.grid, *args)

    def dispatch(self, x, topk_ids):
        """x: [T, H] bf16, topk_ids: [T, topk] int32 (-1 = dropped).  Returns ``(recv_x [W, cap, H], recv_meta [W, cap, 2] = (source token,
        local expert), recv_count [W])`` -- views of the symmetric receive buffers, valid until the next dispatch."""
        T = x.shape[0]
        assert T <= self.T and x.shape[1] == self.H and tuple(topk_ids.shape) == (T, self.topk)
        self.phase += 1
        self._n_tokens = T
        self._ids = topk_ids.contiguous()
        self._run(ep_dispatch, lk.symm_ctx(), x.contiguous(), self._ids, self.recv_x, self.recv_meta, self.send_count, self.send_slot,
                  self.recv_count, self.flags, self.bar, T, self.H, self.topk, self.epr, self.cap, self.phase)
        return self.recv_x, self.recv_meta, self.recv_count[: self.W]

    def combine(self, topk_w, out=None):
        """Expert outputs must be in ``self.y_buf`` (same [W, cap, H] layout as the received rows) on every rank; this call synchronises
        the ranks first.  topk_w: [T, topk] fp32 -> out [T, H] bf16."""
        import torch
        import triton_dist.utils as U
        U.barrier_all_on_stream()
        T = self._n_tokens
        out = torch.empty(T, self.H, dtype=torch.bfloat16, device=self.y_buf.device) if out is None else out
        self._run(ep_combine, lk.symm_ctx(), self.y_buf, self._ids, topk_w.contiguous().float(), self.send_slot, out, T, self.H, self.topk,
                  self.epr, self.cap)
        U.barrier_all_on_stream()          # nobody overwrites y_buf / recv_x while a peer is still pulling
        return out

    def finalize(self):
        import triton_dist.utils as U
        for t in (self.flags, self.recv_count, self.recv_meta, self.y_buf, self.recv_x):
            U.nvshmem_free_tensor_sync(t)
