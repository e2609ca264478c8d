"""Decode-sized linear layer on warp-level tensor cores, in the DSL:  out[B, N] = x[B, K] @ W[N, K]^T  for 9..64 rows.

This is the algorithm of the megakernel's tensor-core LINEAR task (csrc/megakernel.cu ``linear_mma``) written as a stand-alone DSL
kernel, statement for statement, so that the CPU interpreter (with its model of the ``mma.sync`` fragment layout) executes exactly what
the CUDA task does:

* the activations of a K chunk are staged in shared memory in FRAGMENT ORDER -- entry ``((m * C32 + c) * 32 + lane) * 2 + hi`` holds
  row ``m * 16 + g + 8 * hi``, elements ``k = 32 c + 8 tig .. + 7`` (lane = 4 g + tig);
* a lane's 16-byte weight load (row ``n0 + 8 ng + g``, the same 8 k's) IS its B fragments for the two MMAs of the 32-wide chunk: the
  K order inside a chunk is permuted identically for both operands, which a dot product does not notice;
* with fewer than 8 column groups per CTA, ``8 / G`` warps share a group and split the K chunks; partial sums meet in shared memory.

``tests/test_lk_cpu.py`` runs it in the interpreter against ``x @ W^T`` (ragged batch, K split across warps, several passes).
"""
from triton_dist import lk
from triton_dist.lk import ll

THREADS = 256
NW = THREADS // 32
MTMAX = 4                      # up to 64 rows
U = 2                          # weight loads in flight per lane and group (the CUDA task uses 8 / NGW)


def make_linear_mma(KC: int = 256):
    """``KC``: elements of K staged per chunk (multiple of 32; the CUDA task derives it from a 128 KB staging budget)."""
    assert KC % 32 == 0
    FRAG_WORDS = MTMAX * 16 * KC // 2                     # 32-bit words of one staged chunk (bf16 pairs)

    @lk.kernel(block=THREADS)
    def linear_mma(x: ll.ptr[ll.bf16], W: ll.ptr[ll.bf16], out: ll.ptr[ll.bf16], B: ll.i32, K: ll.i32, ldo: ll.i32, n_cnt: ll.i32):
        frag = ll.dyn_shared([FRAG_WORDS], ll.u32)        # staged activations, fragment order, as packed bf16 pairs
        part = ll.dyn_shared([NW * MTMAX * 4 * 32], ll.f32)   # partial sums of the warps that split K
        tid = ll.threadIdx.x
        warp = tid // 32
        lane = tid % 32
        g = lane // 4
        tig = lane % 4
        n0 = ll.blockIdx.x * n_cnt
        mt = (B + 15) // 16
        kvec = K // 8
        groups = n_cnt // 8
        G = 1
        while G < groups and G < NW:
            G = G * 2
        kparts = NW // G                                  # warps that share one column group and split the K chunks
        gi = warp % G
        kp = warp // G
        acc = ll.local([MTMAX * 4], ll.f32)
        pass0 = 0
        while pass0 < groups:
            ng = pass0 + gi
            for e in ll.static_range(MTMAX * 4):
                acc[e] = 0.0
            kc0 = 0
            while kc0 < K:
                kc = min(KC, K - kc0)
                C32 = kc // 32
                ll.syncthreads()                          # the previous chunk has been consumed
                for i in range(tid, mt * 16 * (kc // 8), THREADS):
                    b = i // (kc // 8)
                    kg = i % (kc // 8)
                    v = ll.make_uint4(0, 0, 0, 0)
                    if b < B:
                        v = ll.ld_v4(x + (ll.i64(b) * K + kc0 + kg * 8))
                    m = b // 16
                    r = b % 16
                    w0 = ((((m * C32 + kg // 4) * 32 + (r % 8) * 4 + kg % 4) * 2) + r // 8) * 4
                    frag[w0 + 0] = v.x
                    frag[w0 + 1] = v.y
                    frag[w0 + 2] = v.z
                    frag[w0 + 3] = v.w
                ll.syncthreads()
                c0 = kp * U
                while c0 < C32:
                    for u in ll.static_range(U):
                        if c0 + u < C32:
                            wv = ll.make_uint4(0, 0, 0, 0)
                            if ng < groups:
                                wv = ll.ld_nc_v4(W + ((ll.i64(n0 + ng * 8 + g) * kvec + kc0 // 8 + (c0 + u) * 4 + tig) * 8))
                            for m2 in ll.static_range(MTMAX):
                                if m2 < mt:
                                    f0 = (((m2 * C32 + c0 + u) * 32 + lane) * 2) * 4      # X = row g, Y = row g + 8 (next 4 words)
                                    ll.mma_m16n8k16_bf16(acc, m2 * 4, frag[f0 + 0], frag[f0 + 4], frag[f0 + 1], frag[f0 + 5], wv.x, wv.y)
                                    ll.mma_m16n8k16_bf16(acc, m2 * 4, frag[f0 + 2], frag[f0 + 6], frag[f0 + 3], frag[f0 + 7], wv.z, wv.w)
                    c0 += U * kparts
                kc0 += KC
            if kparts > 1:
                ll.syncthreads()
                if kp > 0:
                    for e2 in ll.static_range(MTMAX * 4):
                        part[(warp * MTMAX * 4 + e2) * 32 + lane] = acc[e2]
                ll.syncthreads()
                if kp == 0:
                    for q in range(1, kparts):
                        w2 = q * G + gi
                        for e3 in ll.static_range(MTMAX * 4):
                            acc[e3] += part[(w2 * MTMAX * 4 + e3) * 32 + lane]
            if kp == 0 and ng < groups:
                col = n0 + ng * 8 + tig * 2
                for m3 in ll.static_range(MTMAX):
                    r0 = m3 * 16 + g
                    if r0 < B:
                        out[ll.i64(r0) * ldo + col] = acc[m3 * 4 + 0]
                        out[ll.i64(r0) * ldo + col + 1] = acc[m3 * 4 + 1]
                    if r0 + 8 < B:
                        out[ll.i64(r0 + 8) * ldo + col] = acc[m3 * 4 + 2]
                        out[ll.i64(r0 + 8) * ldo + col + 1] = acc[m3 * 4 + 3]
            pass0 += G
        ll.syncthreads()

    linear_mma.name = f"lk_linear_mma_kc{KC}"
    return linear_mma


_CACHE = {}


def run_linear_mma(x, W, tile_n: int = 32, KC: int = 256, interpret: bool = False):
    """x: [B, K] bf16 (9 <= B <= 64), W: [N, K] bf16, K % 32 == 0, N % tile_n == 0, tile_n % 8 == 0 -> [B, N] bf16."""
    import torch
    B, K = x.shape
    N = W.shape[0]
    assert 1 <= B <= 64 and K % 32 == 0 and N % tile_n == 0 and tile_n % 8 == 0 and x.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    if KC not in _CACHE:
        _CACHE[KC] = make_linear_mma(KC)
    k = _CACHE[KC]
    out = torch.empty(B, N, dtype=torch.bfloat16, device=x.device)
    args = (x.contiguous(), W.contiguous(), out, B, K, N, tile_n)
    if interpret or not x.is_cuda:
        k.interpret(N // tile_n, *args)
    else:
        k[N // tile_n](*args)
    return out
