"""GEMM + ReduceScatter as ONE kernel written in the Python DSL:  out[Ms, N] = sum_r (A_r[W * Ms, K_r] @ B_r[N, K_r]^T)[my rows].

The second half of the tensor-parallel MLP (csrc/gemm_sm100.cuh mode kRS runs a ring inside the epilogue; reference:
kernels/nvidia/gemm_reduce_scatter.py).  The DSL version uses the other B200-native scheme of this repository (the one of the fused
``moe_reduce_rs`` kernel): every finished tile is ADDED straight into the owner rank's accumulation buffer with 16-byte L2 reductions
over NVLink (``red.global.add.v4.bf16x2`` on a ``symm_at`` address) -- no staging, no ring order -- followed by one release-add of the
owner's arrival counter per tile; the last CTAs of the grid are collectors that acquire ``counter >= W * tiles_per_owner * phase``, copy
the accumulated rows out and clear them for the next call.  bf16 accumulation across ranks (the ring of the hand-written kernel keeps
fp32 partials); one cross-rank barrier between calls.

Runs across processes in the CPU interpreter (emulation heap, pipeline model, CAS-emulated bf16x2 reductions): case ``lk_gemm_rs``.
"""
from triton_dist import lk
from triton_dist.lk import ll

BM, BK, UMMA_K = 128, 64, 16
THREADS = 192


def make_gemm_rs(BN: int = 256, STAGES: int = 4, N_COLLECT: int = 4):
    assert BN % 32 == 0 and 32 <= BN <= 256
    A_BYTES, B_BYTES = BM * BK * 2, BN * BK * 2
    TX_BYTES = A_BYTES + B_BYTES
    TMEM_COLS = max(32, 1 << (BN - 1).bit_length())
    IDESC = ll.make_idesc(1, 1, BM, BN)

    @lk.kernel(block=THREADS)
    def gemm_rs(ctx: ll.SymmCtx, tA: ll.TmaDescriptor, tB: ll.TmaDescriptor, acc: ll.ptr[ll.bf16], counter: ll.ptr[ll.u32],
                out: ll.ptr[ll.bf16], Ms: ll.i32, N: ll.i32, K: ll.i32, num_n: ll.i32, n_tiles: ll.i32, phase: ll.u32):
        tid = ll.threadIdx.x
        W = ll.num_ranks(ctx)
        if ll.blockIdx.x >= n_tiles:
            # ---------------- collector CTA (last in the grid): wait for every rank's tiles of MY rows, copy out, clear ----------------
            cid = ll.blockIdx.x - n_tiles
            tiles_per_owner = ((Ms + BM - 1) // BM) * num_n
            if tid == 0:
                ll.wait_ge(counter, W * tiles_per_owner * phase)
            ll.syncthreads()
            nvec = Ms * (N // 8)
            zero = ll.make_uint4(0, 0, 0, 0)
            for vec in range(cid * THREADS + tid, nvec, N_COLLECT * THREADS):
                ll.st_v4(out + ll.i64(vec) * 8, ll.ld_v4(acc + ll.i64(vec) * 8))
                ll.st_v4(acc + ll.i64(vec) * 8, zero)
            return

        # ---------------- GEMM tile (1-CTA tcgen05 rung); the epilogue reduces into the owner of the rows ----------------
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
        m0 = (tile // num_n) * BM                                # row of the full [W * Ms, N] product
        n0 = (tile % num_n) * BN
        owner = m0 // Ms                                         # Ms % BM == 0: a tile belongs to one owner
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
            dst_base = ll.symm_at(ctx, acc, owner)               # the owner's accumulation buffer (my own when owner == rank)
            for c in ll.static_range(BN // 32):
                ll.tmem_ld_32x32b_x32(tmem + ll.u32((quad * 32) << 16) + c * 32, regs)
                ll.tmem_ld_wait()
                if row < W * Ms:
                    dstc = dst_base + (ll.i64(row - owner * Ms) * N + n0 + c * 32)
                    for j in ll.static_range(4):
                        if n0 + c * 32 + j * 8 < N:
                            v = ll.make_uint4(
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 0]), ll.uint_as_float(regs[8 * j + 1])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 2]), ll.uint_as_float(regs[8 * j + 3])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 4]), ll.uint_as_float(regs[8 * j + 5])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 6]), ll.uint_as_float(regs[8 * j + 7])))
                            ll.red_add_bf16x8(dstc + j * 8, v)
        ll.tc_fence_before()
        ll.syncthreads()                                         # every reduction of this tile has been issued ...
        if tid == 0:
            ll.notify(ctx, counter, owner, 1, op="add")          # ... then one release-add on the owner's arrival counter
        if warp == 1:
            ll.tmem_dealloc(tmem, TMEM_COLS)

    gemm_rs.name = f"lk_gemm_rs_bn{BN}_s{STAGES}_c{N_COLLECT}"
    gemm_rs.n_collect = N_COLLECT
    gemm_rs.bn = BN
    return gemm_rs


class LkGemmRsContext:
    """Symmetric accumulation buffer [max_Ms, N] (zero between calls) + arrival counter; ``phase`` counts calls."""

    def __init__(self, max_Ms: int, N: int, BN: int = 256, STAGES: int = 4, N_COLLECT: int = 4):
        import torch
        import triton_dist.utils as U
        self.W, self.rank, self.max_Ms, self.N = U.world_size(), U.rank(), max_Ms, N
        self.acc = U.nvshmem_create_tensor((max_Ms, N), torch.bfloat16)
        self.counter = U.nvshmem_create_tensor((8,), torch.int32)
        self.acc.zero_()
        self.counter.zero_()
        self.phase = 0
        self.kernel = make_gemm_rs(BN, STAGES, N_COLLECT)
        U.barrier_all_on_stream()

    def finalize(self):
        import triton_dist.utils as U
        U.nvshmem_free_tensor_sync(self.counter)
        U.nvshmem_free_tensor_sync(self.acc)


def run_gemm_rs(ctx: LkGemmRsContext, a, b, out=None):
    """a: [W * Ms, K_local] bf16, b: [N, K_local] bf16 (this rank's K shard of both operands) -> out [Ms, N] bf16 = my rows of the sum."""
    import torch
    import triton_dist.utils as U
    M, K = a.shape
    N = b.shape[0]
    W = ctx.W
    Ms = M // W
    assert M == W * Ms and Ms == ctx.max_Ms and Ms % BM == 0 and N == ctx.N and K % BK == 0 and N % 8 == 0
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    k = ctx.kernel
    out = torch.empty(Ms, N, dtype=torch.bfloat16, device=a.device) if out is None else out
    ctx.phase += 1
    tA = lk.tma_2d(a, BK, BM)
    tB = lk.tma_2d(b, BK, k.bn)
    num_m, num_n = M // BM, (N + k.bn - 1) // k.bn
    n_tiles = num_m * num_n
    args = (lk.symm_ctx(), tA, tB, ctx.acc, ctx.counter, out, Ms, N, K, num_n, n_tiles, ctx.phase)
    if a.is_cuda:
        k[n_tiles + k.n_collect](*args)
    else:
        k.interpret(n_tiles + k.n_collect, *args)
    U.barrier_all_on_stream()          # the next call's reductions must find every owner's buffer cleared
    return out
