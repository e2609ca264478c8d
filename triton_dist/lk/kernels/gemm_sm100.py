"""bf16 GEMM ladder for sm_100a written in the DSL:  C[M, N] = A[M, K] @ B[N, K]^T  (both operands K-major, fp32 accumulation in TMEM).

The reference ships nine "levels" of one tcgen05 GEMM in its DSL (python/little_kernel/benchmark/gemm_sm100/gemm_level{1..9}.py, from a
1-stage 2-SM kernel up to a persistent warp-specialised one).  Two rungs are kept here, following the protocol of the hand-written
kernel of this framework (csrc/gemm_sm100.cuh) so that both share the hardware-validated PTX wrappers:

* ``gemm_1cta``: one CTA per 128 x BN tile, STAGES-deep TMA -> smem ring, warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
  ``tcgen05.mma`` issuer, warps 2..5 = epilogue (``tcgen05.ld`` -> bf16 -> 16-byte global stores).
* ``gemm_2cta``: a 2-CTA cluster per 256 x BN tile (``cta_group::2``): each CTA loads its 128 rows of A and half of B, all bytes land on
  the leader's mbarrier, the leader issues M = 256 MMAs, ``tcgen05.commit`` multicasts to both CTAs, each CTA drains its own TMEM half.
* ``gemm_persistent`` (``make_gemm_persistent``): the ladder's endpoint -- one cluster per SM pair walks the tiles (group-M swizzled
  order), the smem ring runs across tile boundaries, TWO TMEM accumulators let the epilogue of tile i overlap the mainloop of tile
  i + 1 (``tmem_full`` / ``tmem_empty`` mbarriers, one arrival per epilogue warp of the pair).  Same protocol as the hand-written
  kernel; it compiles and its first hardware run is pending (the first two rungs pass on a B200).

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


def make_gemm_persistent(BN: int = 256, STAGES: int = 6, cta_group: int = 2, GROUP_M: int = 8):
    """Persistent warp-specialised kernel: 256 threads = TMA warp, MMA warp, TMEM warp, (idle), 4 epilogue warps."""
    assert BN in (128, 256) and cta_group in (1, 2)
    B_ROWS = BN // cta_group
    A_BYTES, B_BYTES = BM * BK * 2, B_ROWS * BK * 2
    TX_BYTES = A_BYTES + B_BYTES
    TMEM_COLS = 2 * BN                                          # two accumulators
    IDESC = ll.make_idesc(1, 1, BM * cta_group, BN)
    TILE_M = BM * cta_group
    P_THREADS = 256
    assert STAGES * TX_BYTES + 4096 <= 227 * 1024, "smem ring too large: lower STAGES (cta_group 1 with BN = 256 fits 4 stages)"

    @lk.kernel(block=P_THREADS, cluster=(cta_group, 1, 1))
    def gemm(tA: ll.TmaDescriptor, tB: ll.TmaDescriptor, C: ll.ptr[ll.bf16], M: ll.i32, N: ll.i32, K: ll.i32, num_m: ll.i32, num_n: ll.i32):
        ll.align_memory(1024)
        sA = ll.dyn_shared([STAGES, BM * BK], ll.bf16, align=1024)
        sB = ll.dyn_shared([STAGES, B_ROWS * BK], ll.bf16, align=1024)
        full = ll.dyn_shared([STAGES], ll.u64)
        empty = ll.dyn_shared([STAGES], ll.u64)
        tfull = ll.dyn_shared([2], ll.u64)                      # accumulator complete (one tcgen05.commit)
        tempty = ll.dyn_shared([2], ll.u64)                     # accumulator drained (one arrival per epilogue warp of the pair)
        tmem_slot = ll.dyn_shared([4], ll.u32)

        warp = ll.warp_id()
        lane = ll.lane_id()
        cta = ll.cluster_rank() if cta_group == 2 else 0
        n_workers = ll.gridDim.x // cta_group
        worker = ll.blockIdx.x // cta_group
        total = num_m * num_n
        nkb = (K + BK - 1) // BK
        per_band = GROUP_M * num_n

        if warp == 0 and lane == 0:
            ll.prefetch_tensormap(tA)
            ll.prefetch_tensormap(tB)
        if warp == 1 and lane == 0:
            for s in ll.static_range(STAGES):
                ll.mbar_init(full + s, cta_group)
                ll.mbar_init(empty + s, 1)
            for i in ll.static_range(2):
                ll.mbar_init(tfull + i, 1)
                ll.mbar_init(tempty + i, 4 * cta_group)
            ll.fence_barrier_init()
        if warp == 2:
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
            # ---------------- TMA producer: the ring keeps running across tile boundaries ----------------
            if ll.elect_one():
                stage = 0
                phase = 0
                u = worker
                while u < total:
                    band = u // per_band
                    first_m = band * GROUP_M
                    band_m = min(num_m - first_m, GROUP_M)
                    r = u - band * per_band
                    m0 = (first_m + r % band_m) * TILE_M + cta * BM
                    n0 = (r // band_m) * BN + cta * B_ROWS
                    for kb in range(nkb):
                        ll.mbar_wait(empty + stage, phase ^ 1)
                        if cta_group == 1:
                            ll.mbar_arrive_expect_tx(full + stage, TX_BYTES)
                            ll.tma_load_2d(tA, full + stage, sA[stage], kb * BK, m0)
                            ll.tma_load_2d(tB, full + stage, sB[stage], kb * BK, n0)
                        else:
                            if cta == 0:
                                ll.mbar_arrive_expect_tx(full + stage, 2 * TX_BYTES)
                            else:
                                ll.mbar_arrive_cluster(full + stage, 0)
                            ll.tma_load_2d_2sm(tA, full + stage, sA[stage], kb * BK, m0)
                            ll.tma_load_2d_2sm(tB, full + stage, sB[stage], kb * BK, n0)
                        stage += 1
                        if stage == STAGES:
                            stage = 0
                            phase ^= 1
                    u += n_workers
            ll.syncwarp()
        elif warp == 1:
            # ---------------- MMA issuer: alternates between the two TMEM accumulators ----------------
            if cta == 0 and ll.elect_one():
                stage2 = 0
                phase2 = 0
                acc = 0
                acc_phase = 0
                u2 = worker
                while u2 < total:
                    ll.mbar_wait(tempty + acc, acc_phase ^ 1)   # the epilogue has drained this accumulator
                    ll.tc_fence_after()
                    d_tmem = tmem + ll.u32(acc * BN)
                    for kb in range(nkb):
                        ll.mbar_wait(full + stage2, phase2)
                        ll.tc_fence_after()
                        adesc = ll.make_smem_desc_k128(ll.smem_addr(sA[stage2]))
                        bdesc = ll.make_smem_desc_k128(ll.smem_addr(sB[stage2]))
                        for k in ll.static_range(BK // UMMA_K):
                            accumulate = ll.u32(1) if k > 0 else ll.u32(kb > 0)
                            ll.mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, accumulate, cta_group=cta_group)
                        if cta_group == 1:
                            ll.mma_commit(empty + stage2)
                        else:
                            ll.mma_commit_2sm(empty + stage2, 3)
                        stage2 += 1
                        if stage2 == STAGES:
                            stage2 = 0
                            phase2 ^= 1
                    if cta_group == 1:
                        ll.mma_commit(tfull + acc)
                    else:
                        ll.mma_commit_2sm(tfull + acc, 3)
                    acc += 1
                    if acc == 2:
                        acc = 0
                        acc_phase ^= 1
                    u2 += n_workers
            ll.syncwarp()
        elif warp >= 4:
            # ---------------- epilogue warps: drain accumulator `acc3` while the MMA warp fills the other one ----------------
            quad = warp - 4                                     # == warp % 4: the TMEM lanes this warp may read
            acc3 = 0
            acc_phase3 = 0
            regs = ll.local([32], ll.u32)
            u3 = worker
            while u3 < total:
                band3 = u3 // per_band
                first3 = band3 * GROUP_M
                bm3 = min(num_m - first3, GROUP_M)
                r3 = u3 - band3 * per_band
                row = (first3 + r3 % bm3) * TILE_M + cta * BM + quad * 32 + lane
                c0 = (r3 // bm3) * BN
                ll.mbar_wait(tfull + acc3, acc_phase3)
                ll.tc_fence_after()
                for c in ll.static_range(BN // 32):
                    ll.tmem_ld_32x32b_x32(tmem + ll.u32((quad * 32) << 16) + ll.u32(acc3 * BN + c * 32), regs)
                    ll.tmem_ld_wait()
                    if row < M:
                        dst = C + (ll.i64(row) * N + c0 + c * 32)
                        for j in ll.static_range(4):
                            if c0 + c * 32 + j * 8 < N:
                                v = ll.make_uint4(
                                    ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 0]), ll.uint_as_float(regs[8 * j + 1])),
                                    ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 2]), ll.uint_as_float(regs[8 * j + 3])),
                                    ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 4]), ll.uint_as_float(regs[8 * j + 5])),
                                    ll.pack_bf16x2(ll.uint_as_float(regs[8 * j + 6]), ll.uint_as_float(regs[8 * j + 7])))
                                ll.st_v4(dst + j * 8, v)
                ll.tc_fence_before()
                ll.syncwarp()
                if lane == 0:
                    if cta_group == 1:
                        ll.mbar_arrive(tempty + acc3)
                    else:
                        ll.mbar_arrive_cluster(tempty + acc3, 0)  # both CTAs' warps arrive on the LEADER's barrier
                acc3 += 1
                if acc3 == 2:
                    acc3 = 0
                    acc_phase3 ^= 1
                u3 += n_workers
        ll.tc_fence_before()
        if cta_group == 2:
            ll.cluster_sync()
        else:
            ll.syncthreads()
        if warp == 2:
            ll.tmem_dealloc(tmem, TMEM_COLS, cta_group=cta_group)

    gemm.name = f"lk_gemm_persistent_bn{BN}_s{STAGES}_cg{cta_group}" + ("" if GROUP_M == 8 else f"_gm{GROUP_M}")
    gemm.tile = (TILE_M, BN, BK, B_ROWS)
    return gemm


_CACHE = {}


def get_gemm(BN=256, STAGES=4, cta_group=1):
    key = (BN, STAGES, cta_group)
    if key not in _CACHE:
        _CACHE[key] = make_gemm(*key)
    return _CACHE[key]


def run_gemm_persistent(a, b, out=None, BN: int = 256, STAGES: int = 6, cta_group: int = 2, num_sms: int = 148, GROUP_M: int = 8):
    """The persistent rung: grid = one cluster per SM (pair); same operand contract as :func:`run_gemm`."""
    import torch
    M, K = a.shape
    N = b.shape[0]
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and b.shape[1] == K and K % BK == 0 and N % 8 == 0
    key = ("persistent", BN, STAGES, cta_group, GROUP_M)
    if key not in _CACHE:
        _CACHE[key] = make_gemm_persistent(BN, STAGES, cta_group, GROUP_M)
    k = _CACHE[key]
    tile_m, _, _, b_rows = k.tile
    out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16) if out is None else out
    tA = lk.tma_2d(a, BK, BM)
    tB = lk.tma_2d(b, BK, b_rows)
    num_m, num_n = (M + tile_m - 1) // tile_m, (N + BN - 1) // BN
    workers = max(1, min(num_sms // cta_group, num_m * num_n))
    if a.is_cuda:
        k[workers * cta_group](tA, tB, out, M, N, K, num_m, num_n)
    else:
        k.interpret(workers * cta_group, tA, tB, out, M, N, K, num_m, num_n)
    return out


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
    if a.is_cuda:
        k[grid](tA, tB, out, M, N, K)
    else:                       # CPU tensors: the interpreter's functional pipeline model (a few tiles at most)
        k.interpret(grid, tA, tB, out, M, N, K)
    return out


# The ladder as a table (reference: gemm_level1..9.py + test_all_levels.py): each level adds ONE idea to the previous one; all of them are
# configurations of the two factories above, so every level shares the hardware-validated PTX wrappers and runs in the CPU pipeline model.
LEVELS = {
    1: ("1 CTA per 128 x 128 tile, single-stage: TMA -> tcgen05.mma -> tcgen05.ld, nothing overlaps", dict(fn="tile", BN=128, STAGES=1, cta_group=1)),
    2: ("2-stage smem ring: the next TMA load overlaps the current MMAs", dict(fn="tile", BN=128, STAGES=2, cta_group=1)),
    3: ("4-stage ring: loads run a full pipeline depth ahead", dict(fn="tile", BN=128, STAGES=4, cta_group=1)),
    4: ("256-wide tiles: half the A traffic per FLOP, the whole 512-column TMEM budget of one accumulator", dict(fn="tile", BN=256, STAGES=4, cta_group=1)),
    5: ("cta_group::2: a CTA pair shares one 256 x 256 tile, B is split across the pair, commits are multicast", dict(fn="tile", BN=256, STAGES=4, cta_group=2)),
    6: ("persistent 1-CTA workers: the smem ring runs across tile boundaries", dict(fn="persistent", BN=128, STAGES=4, cta_group=1, GROUP_M=1)),
    7: ("persistent + two TMEM accumulators: the epilogue of tile i overlaps the mainloop of tile i + 1", dict(fn="persistent", BN=256, STAGES=4, cta_group=1, GROUP_M=1)),
    8: ("persistent CTA pairs, 6-stage ring, row-major tile order", dict(fn="persistent", BN=256, STAGES=6, cta_group=2, GROUP_M=1)),
    9: ("level 8 with the group-M swizzled tile order (L2 reuse of B across neighbouring M tiles): the hand-written kernel's schedule",
        dict(fn="persistent", BN=256, STAGES=6, cta_group=2, GROUP_M=8)),
}


def run_level(level: int, a, b, out=None, num_sms: int = 148):
    """Run one level of the ladder (same operand contract as :func:`run_gemm`)."""
    _, cfg = LEVELS[level]
    if cfg["fn"] == "tile":
        return run_gemm(a, b, out, cfg["BN"], cfg["STAGES"], cfg["cta_group"])
    return run_gemm_persistent(a, b, out, cfg["BN"], cfg["STAGES"], cfg["cta_group"], num_sms=num_sms, GROUP_M=cfg["GROUP_M"])


def test_all_levels(M: int = 4096, N: int = 4096, K: int = 4096, device: str = "cuda", levels=None, num_sms: int = 148):
    """Check every level against fp32 and (on a GPU) time it -- the reference's ``test_all_levels.py``.  Returns ``{level: (max error, ms)}``."""
    import torch
    torch.manual_seed(0)
    a = (torch.randn(M, K, device=device) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=device) * 0.5).bfloat16()
    ref = a.float() @ b.float().t()
    res = {}
    for lv in (levels or sorted(LEVELS)):
        c = run_level(lv, a, b, num_sms=num_sms)
        ms = 0.0
        if a.is_cuda:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                run_level(lv, a, b, out=c, num_sms=num_sms)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
        res[lv] = ((c.float() - ref).abs().max().item(), ms)
    return res


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
