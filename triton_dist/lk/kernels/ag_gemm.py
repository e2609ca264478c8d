"""AllGather + GEMM as ONE kernel written in the Python DSL:  C[W * Ms, N] = all_gather(A_shard)[W * Ms, K] @ B[N, K]^T.

The flagship fused op of the framework (csrc/gemm_sm100.cuh mode kAG; reference: kernels/nvidia/allgather_gemm.py) restated in the DSL
to show that one Python function can hold both halves of a compute-communication kernel:

* CTAs ``0 .. N_COMM-1`` are communication CTAs: each takes a slice of this rank's A rows and stores it (16-byte vectors over NVLink,
  ``ll.symm_at``) into slot ``rank`` of EVERY rank's symmetric workspace, then adds 1 to ``flags[rank]`` on that peer with release
  semantics (``ll.notify(..., op="add")``) -- one arrival per (source, comm CTA);
* the remaining CTAs are tcgen05 GEMM tiles (the 1-CTA rung of the ladder): the TMA-issuing thread of a tile acquires
  ``flags[source of its rows] >= N_COMM * phase`` (``ll.wait_ge``), fences the async proxy, and only then loads A tiles from the
  workspace -- tiles of rows that have arrived start while later shards are still in flight.

Flags are monotone (``phase`` = call number), the workspace is double-buffered by call parity; callers put one cross-rank barrier
between calls (``run_ag_gemm`` does) so nobody runs two calls ahead of a peer that is still reading.

The same source runs in the CPU interpreter: across processes on the emulation backend, with the functional pipeline model standing in
for TMA / TMEM / tcgen05 (``tests/dist_worker.py`` case ``lk_ag_gemm``).  Comm CTAs come first in the grid so they are always resident
before any tile that waits for them.
"""
from triton_dist import lk
from triton_dist.lk import ll

BM, BK, UMMA_K = 128, 64, 16
THREADS = 192


def make_ag_gemm(BN: int = 256, STAGES: int = 4, N_COMM: int = 4):
    assert BN % 32 == 0 and 32 <= BN <= 256
    A_BYTES, B_BYTES = BM * BK * 2, BN * BK * 2
    TX_BYTES = A_BYTES + B_BYTES
    TMEM_COLS = max(32, 1 << (BN - 1).bit_length())
    IDESC = ll.make_idesc(1, 1, BM, BN)

    @lk.kernel(block=THREADS)
    def ag_gemm(ctx: ll.SymmCtx, tWs: ll.TmaDescriptor, tB: ll.TmaDescriptor, a_local: ll.ptr[ll.bf16], ws: ll.ptr[ll.bf16],
                flags: ll.ptr[ll.u32], C: ll.ptr[ll.bf16], Ms: ll.i32, N: ll.i32, K: ll.i32, num_n: ll.i32, phase: ll.u32):
        tid = ll.threadIdx.x
        me = ll.rank(ctx)
        W = ll.num_ranks(ctx)
        par = ll.i32(phase & 1)
        if ll.blockIdx.x < N_COMM:
            # ---------------- communication CTA: my rows [r0, r1) -> slot `me` of every rank's workspace ----------------
            rows_per = (Ms + N_COMM - 1) // N_COMM
            r0 = ll.blockIdx.x * rows_per
            r1 = min(Ms, r0 + rows_per)
            nvec = (r1 - r0) * (K // 8)
            src = a_local + ll.i64(r0) * K
            for q in range(W):
                peer = (me + q) % W                              # every rank starts with a different destination
                dst = ll.symm_at(ctx, ws, peer) + ((ll.i64(par) * W + me) * Ms + r0) * K
                for vec in range(tid, nvec, THREADS):
                    ll.st_v4(dst + ll.i64(vec) * 8, ll.ld_v4(src + ll.i64(vec) * 8))
                ll.syncthreads()                                 # the block's stores are issued ...
                if tid == 0:
                    ll.notify(ctx, flags + me, peer, 1, op="add")   # ... then one release-add per (source, comm CTA)
            return

        # ---------------- GEMM tile (1-CTA tcgen05 rung) over the gathered rows ----------------
        ll.align_memory(1024)
        sA = ll.dyn_shared([STAGES, BM * BK], ll.bf16, align=1024)
        sB = ll.dyn_shared([STAGES, BN * BK], ll.bf16, align=1024)
        full = ll.dyn_shared([STAGES], ll.u64)
        empty = ll.dyn_shared([STAGES], ll.u64)
        acc_bar = ll.dyn_shared([1], ll.u64)
        tmem_slot = ll.dyn_shared([4], ll.u32)
        warp = ll.warp_id()
        lane = ll.lane_id()
        tile = ll.blockIdx.x - N_COMM
        m0 = (tile // num_n) * BM                                # row in the gathered [W * Ms, K] matrix
        n0 = (tile % num_n) * BN
        nkb = (K + BK - 1) // BK
        if warp == 0 and lane == 0:
            ll.prefetch_tensormap(tWs)
            ll.prefetch_tensormap(tB)
        if warp == 1 and lane == 0:
            for s in ll.static_range(STAGES):
                ll.mbar_init(full + s, 1)
                ll.mbar_init(empty + s, 1)
            ll.mbar_init(acc_bar, 1)
            ll.fence_barrier_init()
        if warp == 1:
            ll.tmem_alloc(tmem_slot, TMEM_COLS)
            ll.tmem_relinquish()
        ll.tc_fence_before()
        ll.syncthreads()
        ll.tc_fence_after()
        tmem = tmem_slot[0]

        if warp == 0:
            if ll.elect_one():
                # the rows of this tile come from one source (Ms % BM == 0) or from two neighbours: wait for every source it touches
                s_first = m0 // Ms
                s_last = min(W * Ms - 1, m0 + BM - 1) // Ms
                for src_rank in range(s_first, s_last + 1):
                    ll.wait_ge(flags + src_rank, N_COMM * phase)
                ll.fence_proxy_async()                           # the shard was written by generic-proxy stores, TMA reads it
                for kb in range(nkb):
                    s = kb % STAGES
                    ph = (kb // STAGES) & 1
                    ll.mbar_wait(empty + s, ph ^ 1)
                    ll.mbar_arrive_expect_tx(full + s, TX_BYTES)
                    ll.tma_load_2d(tWs, full + s, sA[s], kb * BK, par * W * Ms + m0)
                    ll.tma_load_2d(tB, full + s, sB[s], kb * BK, n0)
            ll.syncwarp()
        elif warp == 1:
            if ll.elect_one():
                for kb in range(nkb):
                    s = kb % STAGES
                    ph = (kb // STAGES) & 1
                    ll.mbar_wait(full + s, ph)
                    ll.tc_fence_after()
                    adesc = ll.make_smem_desc_k128(ll.smem_addr(sA[s]))
                    bdesc = ll.make_smem_desc_k128(ll.smem_addr(sB[s]))
                    for k in ll.static_range(BK // UMMA_K):
                        accumulate = ll.u32(1) if k > 0 else ll.u32(kb > 0)
                        ll.mma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, accumulate)
                    ll.mma_commit(empty + s)
                ll.mma_commit(acc_bar)
            ll.syncwarp()
        else:
            ll.mbar_wait(acc_bar, 0)
            ll.tc_fence_after()
            quad = warp % 4
            row = m0 + quad * 32 + lane
            regs = ll.local([32], ll.u32)
            for c in ll.static_range(BN // 32):
                ll.tmem_ld_32x32b_x32(tmem + ll.u32((quad * 32) << 16) + c * 32, regs)
                ll.tmem_ld_wait()
                if row < W * Ms:
                    dstc = C + (ll.i64(row) * N + n0 + c * 32)
                    for j in ll.static_range(4):
                        if n0 + c * 32 + j * 8 < N:
                            v = ll.make_uint4(
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 0]), ll.uint_as_float(regs[8 * j + 1])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 2]), ll.uint_as_float(regs[8 * j + 3])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 4]), ll.uint_as_float(regs[8 * j + 5])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 6]), ll.uint_as_float(regs[8 * j + 7])))
                            ll.st_v4(dstc + j * 8, v)
        ll.tc_fence_before()
        ll.syncthreads()
        if warp == 1:
            ll.tmem_dealloc(tmem, TMEM_COLS)

    ag_gemm.name = f"lk_ag_gemm_bn{BN}_s{STAGES}_c{N_COMM}"
    ag_gemm.n_comm = N_COMM
    ag_gemm.bn = BN
    return ag_gemm


class LkAgGemmContext:
    """Symmetric workspace [2, W * max_Ms, K] + flags [W] for the DSL kernel; ``phase`` counts calls."""

    def __init__(self, max_Ms: int, K: int, BN: int = 256, STAGES: int = 4, N_COMM: int = 4):
        import torch
        import triton_dist.utils as U
        self.W, self.rank = U.world_size(), U.rank()
        self.max_Ms, self.K = max_Ms, K
        self.ws = U.nvshmem_create_tensor((2 * self.W * max_Ms, K), torch.bfloat16)
        self.flags = U.nvshmem_create_tensor((max(self.W, 8),), torch.int32)
        self.flags.zero_()
        self.phase = 0
        self.kernel = make_ag_gemm(BN, STAGES, N_COMM)
        U.barrier_all_on_stream()

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.flags)
        U.nvshmem_free_tensor_sync(self.ws)


def run_ag_gemm(ctx: LkAgGemmContext, a_shard, b, out=None):
    """a_shard: [Ms, K] bf16 (this rank's rows, Ms == ctx.max_Ms), b: [N, K] bf16 -> [W * Ms, N] bf16."""
    import torch
    import triton_dist.utils as U
    Ms, K = a_shard.shape
    N = b.shape[0]
    assert Ms == ctx.max_Ms and K == ctx.K and K % BK == 0 and N % 8 == 0 and a_shard.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    k = ctx.kernel
    W = ctx.W
    out = torch.empty(W * Ms, N, dtype=torch.bfloat16, device=a_shard.device) if out is None else out
    ctx.phase += 1
    tWs = lk.tma_2d(ctx.ws, BK, BM)
    tB = lk.tma_2d(b, BK, k.bn)
    num_m, num_n = (W * Ms + BM - 1) // BM, (N + k.bn - 1) // k.bn
    grid = k.n_comm + num_m * num_n
    args = (lk.symm_ctx(), tWs, tB, a_shard.contiguous(), ctx.ws, ctx.flags, out, Ms, N, K, num_n, ctx.phase)
    if a_shard.is_cuda:
        k[grid](*args)
    else:
        k.interpret(grid, *args)
    U.barrier_all_on_stream()          # nobody starts call i + 1 while a peer still reads the workspace of call i - 1's parity
    return out
