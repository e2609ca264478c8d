"""bf16 GEMM ladder for sm_100a written in the DSL:  C[M, N] = A[M, K] @ B[N, K]^T  (both operands K-major, fp32 accumulation in TMEM).

The reference ships nine "levels" of one tcgen05 GEMM in its DSL (python/little_kernel/benchmark/gemm_sm100/gemm_level{1..9}.py, from a
1-stage 2-SM kernel up to a persistent warp-specialised one).  Two rungs are kept here, following the protocol of the hand-written
kernel of this framework (csrc/gemm_sm100.cuh) so that both share the hardware-validated PTX wrappers:

* ``gemm_1cta``: one CTA per 128 x BN tile, STAGES-deep TMA -> smem ring, warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
  ``tcgen05.mma`` issuer, warps 2..5 = epilogue (``tcgen05.ld`` -> bf16 -> 16-byte global stores).
* ``gemm_2cta``: a 2-CTA cluster per 256 x BN tile (``cta_group::2``): each CTA loads its 128 rows of A and half of B, all bytes land on
  the leader's mbarrier, the leader issues M = 256 MMAs, ``tcgen05.commit`` multicasts to both CTAs, each CTA drains its own TMEM half.

``run_gemm(a, b, variant=)`` builds the tensor maps and launches; ``python -m triton_dist.lk.kernels.gemm_sm100`` on a B200 checks both
against fp32 and prints their throughput next to the hand-written kernel.
"""
from __future__ import annotations

from triton_dist import lk
from triton_dist.lk import ll

BM, BK, UMMA_K = 128, 64, 16
THREADS = 192                       # warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: epilogue (TMEM lane quadrant = warp % 4)


def make_gemm(BN: int = 256, STAGES: int = 4, cta_group: int = 1):
    """Kernel factory: tile width, pipeline depth and CTA-pair mode are compile-time constants of the generated kernel."""
    assert BN % 32 == 0 and 32 <= BN <= 256 and cta_group in (1, 2)
    B_ROWS = BN // cta_group                                    # rows of B this CTA loads
    A_BYTES, B_BYTES = BM * BK * 2, B_ROWS * BK * 2
    TX_BYTES = A_BYTES + B_BYTES
    TMEM_COLS = max(32, 1 << (BN - 1).bit_length())             # power of two >= BN
    IDESC = ll.make_idesc(1, 1, BM * cta_group, BN)            # bf16 x bf16 -> fp32, M = 128 (or 256 across the pair)
    TILE_M = BM * cta_group

    @lk.kernel(block=THREADS, cluster=(cta_group, 1, 1))
    def gemm(tA: ll.TmaDescriptor, tB: ll.TmaDescriptor, C: ll.ptr[ll.bf16], M: ll.i32, N: ll.i32, K: ll.i32):
        ll.align_memory(1024)
        sA = ll.dyn_shared([STAGES, BM * BK], ll.bf16, align=1024)
        sB = ll.dyn_shared([STAGES, B_ROWS * BK], ll.bf16, align=1024)
        full = ll.dyn_shared([STAGES], ll.u64)
        empty = ll.dyn_shared([STAGES], ll.u64)
        acc_bar = ll.dyn_shared([1], ll.u64)
        tmem_slot = ll.dyn_shared([4], ll.u32)

        warp = ll.warp_id()
        lane = ll.lane_id()
        cta = ll.cluster_rank() if cta_group == 2 else 0
        tile_n = ll.blockIdx.x // cta_group                     # cluster index along x
        m0 = ll.blockIdx.y * TILE_M + cta * BM                  # my 128 rows of A / C
        n0 = tile_n * BN
        nkb = (K + BK - 1) // BK

        if warp == 0 and lane == 0:
            ll.prefetch_tensormap(tA)
            ll.prefetch_tensormap(tB)
        if warp == 1 and lane == 0:
            for s in ll.static_range(STAGES):
                ll.mbar_init(full + s, cta_group)               # one producer arrival per CTA of the pair (+ tx bytes)
                ll.mbar_init(empty + s, 1)                      # one tcgen05.commit
            ll.mbar_init(acc_bar, 1)
            ll.fence_barrier_init()
        if warp == 1:
            ll.tmem_alloc(tmem_slot, TMEM_COLS, cta_group=cta_group)
            ll.tmem_relinquish(cta_group=cta_group)
        ll.tc_fence_before()
        if cta_group == 2:
            ll.cluster_sync()
        else:
            ll.syncthreads()
        ll.tc_fence_after()
        tmem = tmem_slot[0]

        if warp == 0:
            # ---------------- TMA producer (one elected thread) ----------------
            if ll.elect_one():
                for kb in range(nkb):
                    s = kb % STAGES
                    ph = (kb // STAGES) & 1
                    ll.mbar_wait(empty + s, ph ^ 1)             # passes at once on a fresh barrier (parity trick)
                    if cta_group == 1:
                        ll.mbar_arrive_expect_tx(full + s, TX_BYTES)
                        ll.tma_load_2d(tA, full + s, sA[s], kb * BK, m0)
                        ll.tma_load_2d(tB, full + s, sB[s], kb * BK, n0)
                    else:
                        if cta == 0:                            # both CTAs land their bytes on the leader's barrier
                            ll.mbar_arrive_expect_tx(full + s, 2 * TX_BYTES)
                        else:
                            ll.mbar_arrive_cluster(full + s, 0)
                        ll.tma_load_2d_2sm(tA, full + s, sA[s], kb * BK, m0)
                        ll.tma_load_2d_2sm(tB, full + s, sB[s], kb * BK, n0 + cta * B_ROWS)
            ll.syncwarp()
        elif warp == 1:
            # ---------------- MMA issuer (leader CTA, one thread) ----------------
            if cta == 0 and ll.elect_one():
                for kb in range(nkb):
                    s = kb % STAGES
                    ph = (kb // STAGES) & 1
                    ll.mbar_wait(full + s, ph)
                    ll.tc_fence_after()
                    adesc = ll.make_smem_desc_k128(ll.smem_addr(sA[s]))
                    bdesc = ll.make_smem_desc_k128(ll.smem_addr(sB[s]))
                    for k in ll.static_range(BK // UMMA_K):
                        # +32 bytes along K inside the 128-byte swizzle atom = +2 in the 16-byte address field
                        accumulate = ll.u32(1) if k > 0 else ll.u32(kb > 0)
                        ll.mma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, accumulate, cta_group=cta_group)
                    if cta_group == 1:
                        ll.mma_commit(empty + s)                # smem stage free once these MMAs have read it
                    else:
                        ll.mma_commit_2sm(empty + s, 3)
                if cta_group == 1:
                    ll.mma_commit(acc_bar)                      # accumulator complete
                else:
                    ll.mma_commit_2sm(acc_bar, 3)
            ll.syncwarp()
        else:
            # ---------------- epilogue: TMEM -> registers -> bf16 -> global ----------------
            ll.mbar_wait(acc_bar, 0)
            ll.tc_fence_after()
            quad = warp % 4                                     # the TMEM lanes this warp may read
            row = m0 + quad * 32 + lane
            regs = ll.local([32], ll.u32)
            for c in ll.static_range(BN // 32):
                ll.tmem_ld_32x32b_x32(tmem + ll.u32((quad * 32) << 16) + c * 32, regs)
                ll.tmem_ld_wait()
                if row < M:
                    dst = C + (ll.i64(row) * N + n0 + c * 32)
                    for j in ll.static_range(4):
                        if n0 + c * 32 + j * 8 < N:
                            v = ll.make_uint4(
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 0]), ll.uint_as_float(regs[8 * j + 1])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 2]), ll.uint_as_float(regs[8 * j + 3])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 4]), ll.uint_as_float(regs[8 * j + 5])),
                                ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 6]), ll.uint_as_float(regs[8 * j + 7])))
                            ll.st_v4(dst + j * 8, v)
        ll.tc_fence_before()
        if cta_group == 2:
            ll.cluster_sync()
        else:
            ll.syncthreads()
        if warp == 1:
            ll.tmem_dealloc(tmem, TMEM_COLS, cta_group=cta_group)

    gemm.name = f"lk_gemm_bn{BN}_s{STAGES}_cg{cta_group}"
    gemm.tile = (TILE_M, BN, BK, B_ROWS)
    return gemm


_CACHE = {}


def get_gemm(BN=256, STAGES=4, cta_group=1):
    key = (BN, STAGES, cta_group)
    if key not in _CACHE:
        _CACHE[key] = make_gemm(*key)
    return _CACHE[key]


def run_gemm(a, b, out=None, BN: int = 256, STAGES: int = 4, cta_group: int = 1):
    """a [M, K], b [N, K] bf16 row-major (K % 64 == 0, N % 8 == 0) -> [M, N] bf16."""
    import torch
    M, K = a.shape
    N = b.shape[0]
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and b.shape[1] == K and K % BK == 0 and N % 8 == 0
    k = get_gemm(BN, STAGES, cta_group)
    tile_m, _, _, b_rows = k.tile
    out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16) if out is None else out
    tA = lk.tma_2d(a, BK, BM)                   # box: 64 elements (128 B, SWIZZLE_128B) x 128 rows
    tB = lk.tma_2d(b, BK, b_rows)
    grid = ((N + BN - 1) // BN * cta_group, (M + tile_m - 1) // tile_m)
    k[grid](tA, tB, out, M, N, K)
    return out


if __name__ == "__main__":
    import torch
    torch.manual_seed(0)
    M, N, K = 4096, 4096, 4096
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    ref = a.float() @ b.float().t()
    for cg in (1, 2):
        c = run_gemm(a, b, cta_group=cg)
        torch.cuda.synchronize()
        err = (c.float() - ref).abs().max().item()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            run_gemm(a, b, out=c, cta_group=cg)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 20
        print(f"lk gemm cta_group={cg}: max|err|={err:.3f}  {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s")
