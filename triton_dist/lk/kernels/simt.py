"""Small DSL kernels: every one runs on the GPU *and* in the CPU interpreter (tests/test_lk_cpu.py checks the interpreter against
PyTorch; tests/test_zz_lk_gpu.py checks the compiled code).  They double as the DSL tour of tutorials/10_kernel_dsl.py."""
from __future__ import annotations

from triton_dist import lk
from triton_dist.lk import ll, shmem

BLOCK = 128
NWARPS = BLOCK // 32


@lk.kernel(block=BLOCK)
def saxpy(x: ll.ptr[ll.f32], y: ll.ptr[ll.f32], a: ll.f32, n: ll.i32):
    i = ll.blockIdx.x * ll.blockDim.x + ll.threadIdx.x
    if i < n:
        y[i] = a * x[i] + y[i]


def warp_sum(v):
    """Butterfly reduction; a plain Python helper -> inlined ``__device__`` function specialised on the type of ``v``."""
    for off in ll.static_range(4, -1, -1):
        v = v + ll.shfl_xor(v, 1 << off)
    return v


def warp_max(v):
    for off in ll.static_range(4, -1, -1):
        v = max(v, ll.shfl_xor(v, 1 << off))
    return v


@lk.kernel(block=BLOCK)
def block_sum(x: ll.ptr[ll.bf16], out: ll.ptr[ll.f32], n: ll.i32):
    """Grid-stride sum of a bf16 vector: registers -> warp shuffle -> shared memory -> one atomic per block."""
    part = ll.shared([NWARPS], ll.f32)
    tid = ll.threadIdx.x
    acc: ll.f32 = 0.0
    for i in range(ll.blockIdx.x * BLOCK + tid, n, ll.gridDim.x * BLOCK):
        acc += x[i]
    acc = warp_sum(acc)
    if tid % 32 == 0:
        part[tid // 32] = acc
    ll.syncthreads()
    if tid == 0:
        s: ll.f32 = 0.0
        for w in ll.static_range(NWARPS):
            s += part[w]
        ll.atomic_add(out, s)


@lk.kernel(block=BLOCK)
def softmax_rows(x: ll.ptr[ll.f32], y: ll.ptr[ll.f32], cols: ll.i32):
    """One block per row, two passes (max, then exp-sum) with block-wide reductions through shared memory."""
    red = ll.shared([NWARPS], ll.f32)
    tid = ll.threadIdx.x
    row = x + ll.i64(ll.blockIdx.x) * cols
    out = y + ll.i64(ll.blockIdx.x) * cols
    m: ll.f32 = -3.0e38
    for c in range(tid, cols, BLOCK):
        m = max(m, row[c])
    m = warp_max(m)
    if tid % 32 == 0:
        red[tid // 32] = m
    ll.syncthreads()
    m = red[0]
    for w in ll.static_range(1, NWARPS):
        m = max(m, red[w])
    ll.syncthreads()
    s: ll.f32 = 0.0
    for c in range(tid, cols, BLOCK):
        s += ll.exp(row[c] - m)
    s = warp_sum(s)
    if tid % 32 == 0:
        red[tid // 32] = s
    ll.syncthreads()
    s = 0.0
    for w in ll.static_range(NWARPS):
        s += red[w]
    inv = 1.0 / s
    for c in range(tid, cols, BLOCK):
        out[c] = ll.exp(row[c] - m) * inv


@lk.kernel(block=BLOCK)
def histogram(ids: ll.ptr[ll.i32], counts: ll.ptr[ll.i32], n: ll.i32, nbins: ll.i32):
    """Shared-memory privatised histogram (what MoE routing does per expert): dynamic shared memory + atomics."""
    local_counts = ll.dyn_shared([1024], ll.i32)
    tid = ll.threadIdx.x
    for b in range(tid, nbins, BLOCK):
        local_counts[b] = 0
    ll.syncthreads()
    for i in range(ll.blockIdx.x * BLOCK + tid, n, ll.gridDim.x * BLOCK):
        ll.atomic_add(local_counts + ids[i], 1)
    ll.syncthreads()
    for b in range(tid, nbins, BLOCK):
        v = local_counts[b]
        if v != 0:
            ll.atomic_add(counts + b, v)


# --------------------------------------------------------------------------------------------------------------
# distributed: symmetric-heap kernels (the vocabulary of tutorials 01 / 02, written in the DSL)
# --------------------------------------------------------------------------------------------------------------
@lk.kernel(block=BLOCK)
def ring_shift(ctx: ll.SymmCtx, src: ll.ptr[ll.f32], dst: ll.ptr[ll.f32], flag: ll.ptr[ll.u32], n: ll.i32, phase: ll.u32):
    """Every rank writes its ``src`` into the ``dst`` of rank+1 and raises that rank's flag; then waits for its own flag.
    ``dst`` / ``flag`` live on the symmetric heap; ``phase`` increases per call (no flag reset, CUDA-graph replayable)."""
    me = ll.rank(ctx)
    nxt = (me + 1) % ll.num_ranks(ctx)
    remote = ll.symm_at(ctx, dst, nxt)
    tid = ll.threadIdx.x
    for i in range(tid, n, BLOCK):
        remote[i] = src[i]
    ll.syncthreads()                                  # all stores of the block issued ...
    if tid == 0:
        ll.notify(ctx, flag, nxt, phase)              # ... then the release store of the flag on the peer
    if tid < 32:
        ll.wait(flag, 1, phase)                       # acquire: my predecessor's data is visible after this
    ll.syncthreads()


@lk.kernel(block=BLOCK)
def allgather_push(ctx: ll.SymmCtx, shard: ll.ptr[ll.f32], out: ll.ptr[ll.f32], flags: ll.ptr[ll.u32], n: ll.i32, phase: ll.u32):
    """Full-mesh push all-gather: block b serves peer b -- stores my shard into slot ``rank`` of the peer's ``out`` and raises the
    peer's ``flags[rank]``; every block then waits for all ``world`` flags of this rank."""
    me = ll.rank(ctx)
    world = ll.num_ranks(ctx)
    peer = ll.blockIdx.x
    tid = ll.threadIdx.x
    if peer < world:
        remote = ll.symm_at(ctx, out, peer) + ll.i64(me) * n
        for i in range(tid, n, BLOCK):
            remote[i] = shard[i]
        ll.syncthreads()
        if tid == 0:
            ll.notify(ctx, flags + me, peer, phase)
    if tid < 32:
        ll.wait(flags, world, phase)
    ll.syncthreads()


@lk.kernel(block=128)
def shmem_selftest(ctx: ll.SymmCtx, slots: ll.ptr[ll.u32], epoch: ll.ptr[ll.u32], fc_dst: ll.ptr[ll.f32], bc_dst: ll.ptr[ll.f32],
                   ring_dst: ll.ptr[ll.f32], sig: ll.ptr[ll.u64], misc: ll.ptr[ll.i32], src: ll.ptr[ll.f32], n: ll.i32, phase: ll.u64):
    """Every family of the OpenSHMEM-style device API (``triton_dist.lk.shmem``) once: fcollect over the world team, broadcast inside
    the even team, put-with-signal around the ring (64-bit signal, CMP_GE wait), a thread-scope ``int_p``, a warp-scope put with an
    unaligned byte count, team translation, quiet, barrier_all -- the DSL twin of the CUDA self-test in tests/dist_worker.py."""
    s = shmem.make_sync(slots, epoch)
    world = shmem.team_world(ctx)
    W = shmem.n_pes(ctx)
    even = shmem.team_split_strided(0, 2, (W + 1) // 2)
    shmem.fcollect_block(ctx, world, s, fc_dst, src, n)
    shmem.broadcast_block(ctx, even, s, bc_dst, src, n, ll.i32(phase % ll.u64((W + 1) // 2)))
    nxt = (shmem.my_pe(ctx) + 1) % W
    shmem.putmem_signal_block(ctx, ring_dst, src, n * 4, sig, phase, shmem.SIGNAL_SET, nxt)
    if ll.threadIdx.x == 0:
        shmem.signal_wait_until(sig, shmem.CMP_GE, phase)
        misc[0] = shmem.team_translate_pe(world, shmem.my_pe(ctx), even)       # my index in the even team or -1
        misc[1] = shmem.team_my_pe(ctx, even)
        shmem.int_p(ctx, misc + 2, shmem.my_pe(ctx) * 10 + ll.i32(phase), nxt)
    if ll.threadIdx.x < 32:
        shmem.putmem_warp(ctx, ring_dst + n, src, 3 * 4 + 2, nxt)              # unaligned size: byte tail
    shmem.quiet()
    shmem.barrier_all_block(ctx, s)


@lk.kernel(block=64)
def shmem_selftest_scopes(ctx: ll.SymmCtx, slots: ll.ptr[ll.u32], epoch: ll.ptr[ll.u32], buf: ll.ptr[ll.i32], got: ll.ptr[ll.i32],
                          sig: ll.ptr[ll.u64], misc: ll.ptr[ll.i32], src: ll.ptr[ll.i32], n: ll.i32, phase: ll.u64):
    """The remaining spellings of ``triton_dist.lk.shmem``: thread- and warp-scope puts / gets, byte-count broadcast and warp-scope
    fcollect, team syncs in every scope, ``remote_ptr``, ``team_pe`` / ``team_n_pes``, signal ADD and the other comparisons.
    ``buf``: symmetric int32 [6 * n + 2 * W * n]; ``got``: local int32 [4 * n]."""
    s = shmem.make_sync(slots, epoch)
    world = shmem.team_world(ctx)
    W = shmem.n_pes(ctx)
    me = shmem.my_pe(ctx)
    nxt = (me + 1) % W
    prv = (me + W - 1) % W
    tid = ll.threadIdx.x
    # region 0: thread-scope put to the successor; region 1: warp-scope nbi put; region 2: block-scope rma put
    if tid == 0:
        shmem.putmem(ctx, buf, src, n * 4, nxt)
    if tid < 32:
        shmem.putmem_nbi_warp(ctx, buf + n, src, n * 4, nxt)
    shmem.putmem_rma_block(ctx, buf + 2 * n, src, n * 4, nxt)
    shmem.fence()
    shmem.sync_all_block(ctx, s)                                        # everybody's three regions are filled
    # gets in the three scopes read the PREDECESSOR's copies of what its predecessor wrote
    if tid == 0:
        shmem.getmem(ctx, got, buf, n * 4, prv)
    if tid < 32:
        shmem.getmem_nbi_warp(ctx, got + n, buf + n, n * 4, prv)
    shmem.getmem_block(ctx, got + 2 * n, buf + 2 * n, n * 4, prv)
    rp = shmem.remote_ptr(ctx, buf, nxt)                                # plain loads through the peer mapping
    if tid < n:
        got[3 * n + tid] = rp[tid]
    # collectives in byte-count / warp spellings
    shmem.broadcastmem_block(ctx, world, s, buf + 3 * n, src, n * 4, ll.i32(phase % ll.u64(W)))
    if tid < 32:
        shmem.fcollect_warp(ctx, world, s, buf + 6 * n, src, n)
        shmem.team_sync_warp(ctx, world, s)
        shmem.barrier_warp(ctx, world, s)
        shmem.sync_all_warp(ctx, s)
        shmem.barrier_all_warp(ctx, s)
    if tid == 0:
        shmem.team_sync(ctx, world, s)
        shmem.barrier(ctx, world, s)
        shmem.sync_all(ctx, s)
        shmem.barrier_all(ctx, s)
        shmem.signal_op(ctx, sig, 1, shmem.SIGNAL_ADD, 0)              # everyone adds 1 on PE 0
        shmem.signal_op(ctx, sig + 1, phase, shmem.SIGNAL_SET, nxt)
        shmem.signal_wait_until(sig + 1, shmem.CMP_EQ, phase)
        shmem.signal_wait_until(sig + 1, shmem.CMP_NE, phase + 1)
        shmem.signal_wait_until(sig + 1, shmem.CMP_GT, phase - 1)
        shmem.signal_wait_until(sig + 1, shmem.CMP_LE, phase)
        shmem.signal_wait_until(sig + 1, shmem.CMP_LT, phase + 1)
        if me == 0:
            shmem.signal_wait_until(sig, shmem.CMP_GE, phase * ll.u64(W))
        misc[0] = shmem.team_pe(world, W - 1)
        misc[1] = shmem.team_n_pes(world)
        misc[2] = shmem.team_translate_pe(world, me, world)
    shmem.team_sync_block(ctx, world, s)
    shmem.barrier_block(ctx, world, s)

