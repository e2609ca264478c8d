"""GEMM + AllReduce as ONE kernel written in the Python DSL, reduced by the NVSwitch:  out[M, N] = sum_r A_r[M, K_r] @ B_r[N, K_r]^T.

The third fused op of tensor parallelism (row-parallel linear with replicated output; csrc/gemm_sm100.cuh mode kAR; reference:
kernels/nvidia/gemm_allreduce.py ``kernel_fused_gemm_allreduce`` :565-604 -- GEMM CTAs + consumer CTAs in one grid, tile-granular flags,
``multimem.ld_reduce`` consumers).  In the DSL:

* GEMM CTAs (the 1-CTA tcgen05 rung of the ladder) store their bf16 tile into this rank's half of a symmetric staging buffer and then
  release-flag ``flags[tile][me]`` on EVERY rank with the call number: a tile is reducible once all W ranks have produced it;
* the last ``N_COMM`` CTAs of the grid are consumers: consumer c takes tiles c, c + N_COMM, ...; one warp acquires the W flags of the
  tile, then every thread reduces 16-byte vectors of the tile THROUGH THE MULTICAST ALIAS of the staging buffer
  (``multimem.ld_reduce.bf16x2``, fp32 accumulation inside the switch) and stores them to ``out`` -- tiles finished early are reduced
  while later tiles are still in their mainloop;
* consumers come last in the grid and GEMM CTAs never wait for them, so the kernel cannot deadlock whatever the residency; the staging
  buffer is double-buffered by call parity and flags carry the call number, so calls need no barrier in between.

Every rank reduces every tile (one-shot: W x the switch traffic of a two-shot, one flag round) -- the right trade for the decode-sized M
this op is used at.  Runs across processes in the CPU interpreter: pipeline model for TMA / TMEM / tcgen05, multicast model for the
switch reduction (``tests/dist_worker.py`` case ``lk_gemm_ar``).
"""
from triton_dist import lk
from triton_dist.lk import ll

BM, BK, UMMA_K = 128, 64, 16
THREADS = 192


def make_gemm_ar(BN: int = 128, STAGES: int = 4, N_COMM: int = 8):
    assert BN % 32 == 0 and 32 <= BN <= 256
    A_BYTES, B_BYTES = BM * BK * 2, BN * BK * 2
    TX_BYTES = A_BYTES + B_BYTES
    TMEM_COLS = max(32, 1 << (BN - 1).bit_length())
    IDESC = ll.make_idesc(1, 1, BM, BN)

    @lk.kernel(block=THREADS)
    def gemm_ar(ctx: ll.SymmCtx, tA: ll.TmaDescriptor, tB: ll.TmaDescriptor, stage: ll.ptr[ll.bf16], flags: ll.ptr[ll.u32],
                out: ll.ptr[ll.bf16], M: ll.i32, N: ll.i32, K: ll.i32, num_n: ll.i32, n_tiles: ll.i32, max_elems: ll.i64, phase: ll.u32):
        tid = ll.threadIdx.x
        W = ll.num_ranks(ctx)
        me = ll.rank(ctx)
        half = stage + ll.i64(ll.i32(phase & 1)) * max_elems       # this call's half of the staging buffer
        if ll.blockIdx.x >= n_tiles:
            # ---------------- consumer CTA: reduce finished tiles through the switch ----------------
            cid = ll.blockIdx.x - n_tiles
            mc = ll.symm_mc(ctx, half)
            for t in range(cid, n_tiles, N_COMM):
                if tid < 32:
                    ll.wait(flags + t * W, W, phase, True)           # all W ranks have staged tile t
                ll.syncthreads()
                tm0 = (t // num_n) * BM
                tn0 = (t % num_n) * BN
                rows = min(BM, M - tm0)
                cols8 = (min(BN, N - tn0) + 7) // 8
                for i in range(tid, rows * cols8, THREADS):
                    off = ll.i64(tm0 + i // cols8) * N + tn0 + (i % cols8) * 8
                    ll.st_v4(out + off, ll.multimem_ld_reduce_bf16x8(mc + off))
                ll.syncthreads()
            return

        # ---------------- GEMM tile (1-CTA tcgen05 rung); the epilogue stages the tile and tells every rank ----------------
        ll.align_memory(1024)
        sA = ll.dyn_shared([STAGES, BM * BK], ll.bf16, align=1024)
        sB = ll.dyn_shared([STAGES, BN * BK], ll.bf16, align=1024)
        full = ll.dyn_shared([STAGES], ll.u64)
        empty = ll.dyn_shared([STAGES], ll.u64)
        acc_bar = ll.dyn_shared([1], ll.u64)
        tmem_slot = ll.dyn_shared([4], ll.u32)
        warp = ll.warp_id()
        lane = ll.lane_id()
        tile = ll.blockIdx.x
        m0 = (tile // num_n) * BM
        n0 = (tile % num_n) * BN
        nkb = (K + BK - 1) // BK
        if warp == 0 and lane == 0:
            ll.prefetch_tensormap(tA)
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
                for kb in range(nkb):
                    s = kb % STAGES
                    ph = (kb // STAGES) & 1
                    ll.mbar_wait(empty + s, ph ^ 1)
                    ll.mbar_arrive_expect_tx(full + s, TX_BYTES)
                    ll.tma_load_2d(tA, full + s, sA[s], kb * BK, m0)
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
                if row < M:
                    dstc = half + (ll.i64(row) * N + n0 + c * 32)
                    for j in ll.static_range(4):
                        if n0 + c * 32 + j * 8 < N:
                            v = ll.make_uint4(
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 0]), ll.uint_as_float(regs[8 * j + 1])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 2]), ll.uint_as_float(regs[8 * j + 3])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 4]), ll.uint_as_float(regs[8 * j + 5])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 6]), ll.uint_as_float(regs[8 * j + 7])))
                            ll.st_v4(dstc + j * 8, v)
        ll.tc_fence_before()
        ll.syncthreads()                                         # the whole tile is staged ...
        if tid == 0:
            for r in range(W):
                ll.notify(ctx, flags + (tile * W + me), (me + r) % W, phase)      # ... then one release flag per rank
        if warp == 1:
            ll.tmem_dealloc(tmem, TMEM_COLS)

    gemm_ar.name = f"lk_gemm_ar_bn{BN}_s{STAGES}_c{N_COMM}"
    gemm_ar.n_comm = N_COMM
    gemm_ar.bn = BN
    return gemm_ar


class LkGemmArContext:
    """Symmetric staging buffer [2, max_M, N] + tile flags; ``phase`` counts calls."""

    def __init__(self, max_M: int, N: int, BN: int = 128, STAGES: int = 4, N_COMM: int = 8):
        import torch
        import triton_dist.utils as U
        self.W, self.rank, self.max_M, self.N = U.world_size(), U.rank(), max_M, N
        self.gpu = U.current_device().type == "cuda"
        if self.gpu and not U.is_nvshmem_multimem_supported():
            raise RuntimeError("the DSL GEMM + AllReduce reduces through the NVLS multicast mapping of the symmetric heap")
        self.stage = U.nvshmem_create_tensor((2, max_M, N), torch.bfloat16)
        n_tiles = ((max_M + BM - 1) // BM) * ((N + BN - 1) // BN)
        self.flags = U.nvshmem_create_tensor((n_tiles * self.W,), torch.int32)
        self.flags.zero_()
        self.phase = 0
        self.kernel = make_gemm_ar(BN, STAGES, N_COMM)
        U.barrier_all_on_stream()

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.flags)
        U.nvshmem_free_tensor_sync(self.stage)


def run_gemm_ar(ctx: LkGemmArContext, a, b, out=None):
    """a: [M, K_local] bf16, b: [N, K_local] bf16 (this rank's K shard) -> out [M, N] bf16 = the sum over ranks, on every rank."""
    import torch
    M, K = a.shape
    N = b.shape[0]
    assert M <= ctx.max_M and N == ctx.N and K % BK == 0 and N % 8 == 0 and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    k = ctx.kernel
    out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device) if out is None else out
    ctx.phase += 1
    tA = lk.tma_2d(a, BK, BM)
    tB = lk.tma_2d(b, BK, k.bn)
    num_m, num_n = (M + BM - 1) // BM, (N + k.bn - 1) // k.bn
    n_tiles = num_m * num_n
    # the staging halves are indexed with the ACTUAL row length N; max_elems is the distance between the two halves
    args = (lk.symm_ctx(), tA, tB, ctx.stage, ctx.flags, out, M, N, K, num_n, n_tiles, ctx.max_M * ctx.N, ctx.phase)
    if a.is_cuda:
        k[n_tiles + k.n_comm](*args)
    else:
        k.interpret(n_tiles + k.n_comm, *args)
    return out
