"""The benchmark kernels.  Conventions: ``iters`` inner repetitions keep the measured region long compared with launch overhead; results
that the compiler could otherwise discard are written out behind a condition that is false at run time (``sink``) or are the checkable
output itself; per-thread cycle counts come from ``ll.clock64()`` around a dependent chain."""
from triton_dist import lk
from triton_dist.lk import ll

BLOCK = 256
ILP = 8

KERNELS = {}


def _reg(k):
    KERNELS[k.name] = k
    return k


# ---- compute ----------------------------------------------------------------------------------------------------------------------------
@_reg
@lk.kernel(block=BLOCK)
def fma_throughput(out: ll.ptr[ll.f32], iters: ll.i32, a: ll.f32, b: ll.f32):
    """ILP independent fp32 FMA chains per thread: 2 * ILP * iters FLOP per thread.  out[gid] = closed form for a = 1, b = 0 (stays x0)."""
    gid = ll.blockIdx.x * BLOCK + ll.threadIdx.x
    x = ll.local([ILP], ll.f32)
    for j in ll.static_range(ILP):
        x[j] = ll.f32(j + 1)
    for it in range(iters):
        for j2 in ll.static_range(ILP):
            x[j2] = ll.fma(x[j2], a, b)
    s: ll.f32 = 0.0
    for j3 in ll.static_range(ILP):
        s += x[j3]
    out[gid] = s


@_reg
@lk.kernel(block=BLOCK)
def sfu_throughput(out: ll.ptr[ll.f32], iters: ll.i32):
    """ILP independent ``ex2.approx`` chains (the SFU op of every softmax): ILP * iters special-function ops per thread."""
    gid = ll.blockIdx.x * BLOCK + ll.threadIdx.x
    x = ll.local([ILP], ll.f32)
    for j in ll.static_range(ILP):
        x[j] = ll.f32(j) * 0.125 - 0.5
    for it in range(iters):
        for j2 in ll.static_range(ILP):
            x[j2] = ll.ex2_approx(x[j2]) - 1.0                     # stays in (-0.5, 0.5): no overflow, no denormals
    s: ll.f32 = 0.0
    for j3 in ll.static_range(ILP):
        s += x[j3]
    out[gid] = s


@_reg
@lk.kernel(block=BLOCK)
def mma_sync_throughput(out: ll.ptr[ll.f32], iters: ll.i32):
    """Back-to-back ``mma.sync.m16n8k16`` (bf16 -> fp32) on 4 independent accumulator tiles per warp: 4 * 4096 * iters FLOP per warp.
    A = B = all ones (0x3F803F80 = two bf16 1.0), so every accumulator equals 16 * iters: the output is checkable."""
    acc = ll.local([16], ll.f32)
    for e in ll.static_range(16):
        acc[e] = 0.0
    one = ll.u32(0x3F803F80)
    for it in range(iters):
        for t in ll.static_range(4):
            ll.mma_m16n8k16_bf16(acc, t * 4, one, one, one, one, one, one)
    s: ll.f32 = 0.0
    for e2 in ll.static_range(16):
        s += acc[e2]
    out[ll.blockIdx.x * BLOCK + ll.threadIdx.x] = s


# ---- latency ----------------------------------------------------------------------------------------------------------------------------
@_reg
@lk.kernel(block=32)
def fma_latency(cycles: ll.ptr[ll.i64], out: ll.ptr[ll.f32], iters: ll.i32, a: ll.f32):
    """One dependent FMA chain per thread: cycles / iters = issue-to-use latency of the fp32 pipe."""
    x: ll.f32 = 1.0
    t0 = ll.clock64()
    for it in range(iters):
        x = ll.fma(x, a, a)
    t1 = ll.clock64()
    out[ll.threadIdx.x] = x
    if ll.threadIdx.x == 0:
        cycles[0] = t1 - t0


@_reg
@lk.kernel(block=32)
def pointer_chase(nxt: ll.ptr[ll.i32], cycles: ll.ptr[ll.i64], end: ll.ptr[ll.i32], steps: ll.i32):
    """Thread 0 walks ``i = nxt[i]``: cycles / steps = load-to-use latency of whatever level the footprint of ``nxt`` fits in
    (a permutation with one cycle over 16 KB: L1; 8 MB: L2; 1 GB: HBM).  ``end[0]`` is the final index (checkable)."""
    if ll.threadIdx.x == 0:
        i = 0
        for w in range(steps):                                    # warm the level being measured
            i = nxt[i]
        i = 0
        t0 = ll.clock64()
        for s in range(steps):
            i = nxt[i]
        t1 = ll.clock64()
        end[0] = i
        cycles[0] = t1 - t0


@_reg
@lk.kernel(block=BLOCK)
def smem_pointer_chase(cycles: ll.ptr[ll.i64], end: ll.ptr[ll.i32], stride: ll.i32, steps: ll.i32):
    """The same walk through shared memory (``i -> (i + stride) % 1024``)."""
    ring = ll.shared([1024], ll.i32)
    tid = ll.threadIdx.x
    for k in range(tid, 1024, BLOCK):
        ring[k] = (k + stride) % 1024
    ll.syncthreads()
    if tid == 0:
        i = 0
        t0 = ll.clock64()
        for s in range(steps):
            i = ring[i]
        t1 = ll.clock64()
        end[0] = i
        cycles[0] = t1 - t0


@_reg
@lk.kernel(block=BLOCK)
def sync_latency(cycles: ll.ptr[ll.i64], counter: ll.ptr[ll.u32], iters: ll.i32):
    """cycles[0]: ``__syncthreads`` round trips of a 256-thread block; cycles[1]: same-address global atomics issued by one thread."""
    tid = ll.threadIdx.x
    t0 = ll.clock64()
    for it in range(iters):
        ll.syncthreads()
    t1 = ll.clock64()
    if tid == 0:
        cycles[0] = t1 - t0
        t2 = ll.clock64()
        for it2 in range(iters):
            ll.atomic_add(counter, 1)
        t3 = ll.clock64()
        cycles[1] = t3 - t2


# ---- memory -----------------------------------------------------------------------------------------------------------------------------
@_reg
@lk.kernel(block=BLOCK)
def global_copy(dst: ll.ptr[ll.u32], src: ll.ptr[ll.u32], nvec: ll.i64, iters: ll.i32):
    """Grid-stride copy with 16-byte vectors, ``iters`` passes: 32 * nvec * iters bytes of traffic.  Footprint decides the level
    (>> 126 MB: HBM; 32 MB: L2)."""
    gid = ll.i64(ll.blockIdx.x) * BLOCK + ll.threadIdx.x
    step = ll.i64(ll.gridDim.x) * BLOCK
    for it in range(iters):
        v = gid
        while v < nvec:
            ll.st_v4(dst + v * 4, ll.ld_v4(src + v * 4))
            v += step


@_reg
@lk.kernel(block=BLOCK)
def global_read(src: ll.ptr[ll.u32], sink: ll.ptr[ll.u32], nvec: ll.i64, iters: ll.i32):
    """Read-only stream (16-byte ``ld.global.nc``), xor-folded so the loads cannot be dropped: 16 * nvec * iters bytes.
    sink[gid] receives the fold (checkable against the host xor)."""
    gid = ll.i64(ll.blockIdx.x) * BLOCK + ll.threadIdx.x
    step = ll.i64(ll.gridDim.x) * BLOCK
    acc: ll.u32 = 0
    for it in range(iters):
        v = gid
        while v < nvec:
            q = ll.ld_nc_v4(src + v * 4)
            acc = acc ^ q.x ^ q.y ^ q.z ^ q.w
            v += step
    sink[gid] = acc


@_reg
@lk.kernel(block=BLOCK)
def smem_stride(cycles: ll.ptr[ll.i64], sink: ll.ptr[ll.u32], stride: ll.i32, iters: ll.i32):
    """Every thread reads ``buf[(tid * stride + it) % 4096]``: stride 1 is conflict-free, stride 2 / 4 / ... 32 serialise 2 / 4 / ... 32
    ways.  cycles[0] = block time for ``iters`` loads per thread; the ratio between strides is the conflict cost."""
    buf = ll.shared([4096], ll.u32)
    tid = ll.threadIdx.x
    for k in range(tid, 4096, BLOCK):
        buf[k] = k
    ll.syncthreads()
    acc: ll.u32 = 0
    t0 = ll.clock64()
    for it in range(iters):
        acc += buf[(tid * stride + it) % 4096]
    t1 = ll.clock64()
    sink[ll.blockIdx.x * BLOCK + tid] = acc
    if tid == 0:
        cycles[ll.blockIdx.x] = t1 - t0


# ---- warp -------------------------------------------------------------------------------------------------------------------------------
@_reg
@lk.kernel(block=BLOCK)
def shuffle_throughput(out: ll.ptr[ll.u32], cycles: ll.ptr[ll.i64], iters: ll.i32):
    """Dependent ``shfl.sync.bfly`` + ``vote.ballot`` chain per warp.  A xor-butterfly over all five offsets is a warp-wide xor reduction:
    after the first round every lane holds the same value (checkable)."""
    tid = ll.threadIdx.x
    v = ll.u32(tid * 2654435761)
    t0 = ll.clock64()
    for it in range(iters):
        for off in ll.static_range(5):
            v = v ^ ll.shfl_xor(v, 1 << off)
        v = v + (ll.ballot(v != 0) & 1)
    t1 = ll.clock64()
    out[ll.blockIdx.x * BLOCK + tid] = v
    if tid == 0:
        cycles[ll.blockIdx.x] = t1 - t0


# ---- SM ---------------------------------------------------------------------------------------------------------------------------------
@_reg
@lk.kernel(block=BLOCK)
def int_ipc(out: ll.ptr[ll.u32], cycles: ll.ptr[ll.i64], iters: ll.i32):
    """ILP independent integer multiply-add chains: instructions per clock per SM = BLOCK * ILP * iters * resident blocks / cycles."""
    tid = ll.threadIdx.x
    x = ll.local([ILP], ll.u32)
    for j in ll.static_range(ILP):
        x[j] = ll.u32(tid + j)
    t0 = ll.clock64()
    for it in range(iters):
        for j2 in ll.static_range(ILP):
            x[j2] = x[j2] * 3 + 1
    t1 = ll.clock64()
    s: ll.u32 = 0
    for j3 in ll.static_range(ILP):
        s += x[j3]
    out[ll.blockIdx.x * BLOCK + tid] = s
    if tid == 0:
        cycles[ll.blockIdx.x] = t1 - t0


@_reg
@lk.kernel(block=BLOCK)
def occupancy_probe(smids: ll.ptr[ll.i32], spin: ll.i32):
    """Every block records the SM it ran on and stays resident for ``spin`` dependent operations: launched with a dynamic shared-memory
    request, the number of blocks per SM id among the first wave is the achieved occupancy for that footprint."""
    x: ll.u32 = ll.threadIdx.x
    for it in range(spin):
        x = x * 3 + 1
    if ll.threadIdx.x == 0:
        smids[ll.blockIdx.x] = ll.i32(ll.smid()) + ll.i32(x & 0)
