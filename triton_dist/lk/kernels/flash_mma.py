"""Prefill attention on warp-level tensor cores, in the DSL:  out = softmax(scale * q k^T [+ causal mask] [soft-capped]) v,  D = 128, GQA.

This is the algorithm of the megakernel's FLASH_ATTN task (csrc/megakernel.cu ``task_flash_attn``; reference: the megakernel's
``flash_attn`` / ``qkv_pack_flash_attn`` tasks, mega_triton_kernel/kernels/flash_attn.py) written as a stand-alone DSL kernel, statement
for statement, so the CPU interpreter -- with its model of the ``mma.sync`` fragment layout -- executes what the CUDA task does:

* a CTA owns ``16 * warps`` query rows of one (batch, q head); warp w owns rows 16 w .. 16 w + 15.  A lane's eight 16-byte loads of its
  two query rows ARE its A fragments: chunk c (32 d's) gives the registers of k-steps 2c and 2c + 1, which permutes the order of d inside
  a chunk -- the K tile is read with the same 16-byte pattern, so both operands see the same permutation and q . k does not notice;
* 64 keys per iteration: K row-major in shared memory (row stride 80 words: conflict-free 16-byte reads), V TRANSPOSED as packed key
  pairs (``Vt[d][key / 2]``, row stride 36 words) so a lane's B fragment of P V is one 32-bit word;
* scores stay in registers: the accumulator layout of S = Q K^T is the A-fragment layout of P V (two adjacent 8-key tiles form one 16-key
  k-step), online softmax in the exp2 domain with row statistics reduced over the four lanes of a quad.

Strides are in elements, so the same kernel reads separate q / k / v tensors or one packed ``[B, S, Hq + 2 Hkv, 128]`` tensor.
``tests/test_lk_cpu.py`` runs it in the interpreter against the fp32 reference (causal and full, GQA, ragged S, soft cap, packed qkv).
"""
from triton_dist import lk
from triton_dist.lk import ll

D = 128
BKV = 64                       # keys per iteration
KS = 80                        # words per K row in shared memory (64 data + 16 pad)
VS = 36                        # words per Vt row (32 key pairs + 4 pad)
LOG2E = 1.4426950408889634


def make_flash_mma(THREADS: int = 256):
    """``THREADS`` / 32 warps -> ``BQ = 16 * warps`` query rows per CTA (the CUDA task uses 256 threads, BQ = 128)."""
    assert THREADS % 32 == 0 and 32 <= THREADS <= 256
    BQ = THREADS // 32 * 16

    @lk.kernel(block=THREADS)
    def flash_mma(q: ll.ptr[ll.bf16], k: ll.ptr[ll.bf16], v: ll.ptr[ll.bf16], out: ll.ptr[ll.bf16], S: ll.i32, Hq: ll.i32, Hkv: ll.i32,
                  q_ts: ll.i32, kv_ts: ll.i32, o_ts: ll.i32, scale: ll.f32, softcap: ll.f32, causal: ll.i32):
        Ksm = ll.dyn_shared([BKV * KS], ll.u32)
        Vt = ll.dyn_shared([D * VS], ll.u32)
        tid = ll.threadIdx.x
        warp = tid // 32
        lane = tid % 32
        g = lane // 4
        tig = lane % 4
        nqb = (S + BQ - 1) // BQ
        qb = ll.blockIdx.x % nqb
        h = (ll.blockIdx.x // nqb) % Hq
        b = ll.blockIdx.x // (nqb * Hq)
        kvh = h // (Hq // Hkv)
        q0 = qb * BQ
        r0 = q0 + warp * 16 + g                           # this lane's rows: r0 and r0 + 8
        qf = ll.local([32], ll.u32)
        for hi in ll.static_range(2):
            row = r0 + 8 * hi
            for c in ll.static_range(4):
                qv = ll.make_uint4(0, 0, 0, 0)
                if row < S:
                    qv = ll.ld_v4(q + ((ll.i64(b) * S + row) * q_ts + h * D + c * 32 + tig * 8))
                # k-step 2c: a0/a1 = x (rows g / g+8), a2/a3 = y;  k-step 2c+1: z, w
                qf[(2 * c) * 4 + hi] = qv.x
                qf[(2 * c) * 4 + 2 + hi] = qv.y
                qf[(2 * c + 1) * 4 + hi] = qv.z
                qf[(2 * c + 1) * 4 + 2 + hi] = qv.w
        o = ll.local([64], ll.f32)
        s = ll.local([32], ll.f32)
        for e in ll.static_range(64):
            o[e] = 0.0
        m0: ll.f32 = -1.0e30
        m1: ll.f32 = -1.0e30
        l0: ll.f32 = 0.0
        l1: ll.f32 = 0.0
        kv_end = S
        if causal != 0:
            kv_end = min(S, q0 + BQ)
        kbase = k + (ll.i64(b) * S * kv_ts + kvh * D)
        vbase = v + (ll.i64(b) * S * kv_ts + kvh * D)
        kv0 = 0
        while kv0 < kv_end:
            ll.syncthreads()                              # the previous tile has been consumed
            for i in range(tid, BKV * 16, THREADS):       # K: 64 keys x 16 chunks of 8 d's, row-major
                key = i // 16
                ch = i % 16
                kq = ll.make_uint4(0, 0, 0, 0)
                if kv0 + key < S:
                    kq = ll.ld_v4(kbase + (ll.i64(kv0 + key) * kv_ts + ch * 8))
                w0 = key * KS + ch * 4
                Ksm[w0 + 0] = kq.x
                Ksm[w0 + 1] = kq.y
                Ksm[w0 + 2] = kq.z
                Ksm[w0 + 3] = kq.w
            for i2 in range(tid, 32 * 16, THREADS):       # V: 32 key pairs x 16 chunks -> Vt[d][pair] = (v[2 pair][d], v[2 pair + 1][d])
                pj = i2 % 32
                ch2 = i2 // 32
                va = ll.make_uint4(0, 0, 0, 0)
                vb = ll.make_uint4(0, 0, 0, 0)
                if kv0 + 2 * pj < S:
                    va = ll.ld_v4(vbase + (ll.i64(kv0 + 2 * pj) * kv_ts + ch2 * 8))
                if kv0 + 2 * pj + 1 < S:
                    vb = ll.ld_v4(vbase + (ll.i64(kv0 + 2 * pj + 1) * kv_ts + ch2 * 8))
                d0 = ch2 * 8
                Vt[(d0 + 0) * VS + pj] = (va.x & 0xFFFF) | ((vb.x & 0xFFFF) << 16)
                Vt[(d0 + 1) * VS + pj] = (va.x >> 16) | (vb.x & 0xFFFF0000)
                Vt[(d0 + 2) * VS + pj] = (va.y & 0xFFFF) | ((vb.y & 0xFFFF) << 16)
                Vt[(d0 + 3) * VS + pj] = (va.y >> 16) | (vb.y & 0xFFFF0000)
                Vt[(d0 + 4) * VS + pj] = (va.z & 0xFFFF) | ((vb.z & 0xFFFF) << 16)
                Vt[(d0 + 5) * VS + pj] = (va.z >> 16) | (vb.z & 0xFFFF0000)
                Vt[(d0 + 6) * VS + pj] = (va.w & 0xFFFF) | ((vb.w & 0xFFFF) << 16)
                Vt[(d0 + 7) * VS + pj] = (va.w >> 16) | (vb.w & 0xFFFF0000)
            ll.syncthreads()
            # ---- S = Q K^T: 8 key tiles of 8 keys, 8 k-steps of 16 d's ----
            for e2 in ll.static_range(32):
                s[e2] = 0.0
            for nt in ll.static_range(8):
                for c2 in ll.static_range(4):
                    kw = (nt * 8 + g) * KS + c2 * 16 + tig * 4
                    ll.mma_m16n8k16_bf16(s, nt * 4, qf[(2 * c2) * 4 + 0], qf[(2 * c2) * 4 + 1], qf[(2 * c2) * 4 + 2], qf[(2 * c2) * 4 + 3],
                                         Ksm[kw + 0], Ksm[kw + 1])
                    ll.mma_m16n8k16_bf16(s, nt * 4, qf[(2 * c2 + 1) * 4 + 0], qf[(2 * c2 + 1) * 4 + 1], qf[(2 * c2 + 1) * 4 + 2],
                                         qf[(2 * c2 + 1) * 4 + 3], Ksm[kw + 2], Ksm[kw + 3])
            # ---- scale, soft cap, mask; row maxima over the quad ----
            mx0: ll.f32 = -1.0e30
            mx1: ll.f32 = -1.0e30
            for nt2 in ll.static_range(8):
                for j in ll.static_range(4):
                    x = s[nt2 * 4 + j] * scale
                    if softcap > 0.0:
                        x = softcap * ll.tanh(x / softcap)
                    x = x * LOG2E
                    kj = kv0 + nt2 * 8 + tig * 2 + (j % 2)
                    qi = r0 + 8 * (j // 2)
                    if kj >= S or (causal != 0 and kj > qi):
                        x = -1.0e30
                    s[nt2 * 4 + j] = x
                    if j < 2:
                        mx0 = max(mx0, x)
                    else:
                        mx1 = max(mx1, x)
            mx0 = max(mx0, ll.shfl_xor(mx0, 1))
            mx0 = max(mx0, ll.shfl_xor(mx0, 2))
            mx1 = max(mx1, ll.shfl_xor(mx1, 1))
            mx1 = max(mx1, ll.shfl_xor(mx1, 2))
            mn0 = max(m0, mx0)
            mn1 = max(m1, mx1)
            corr0 = ll.exp2(m0 - mn0)
            corr1 = ll.exp2(m1 - mn1)
            m0 = mn0
            m1 = mn1
            rs0: ll.f32 = 0.0
            rs1: ll.f32 = 0.0
            for nt3 in ll.static_range(8):
                for j2 in ll.static_range(4):
                    pv: ll.f32 = 0.0
                    if s[nt3 * 4 + j2] > -1.0e29:          # masked scores contribute exactly 0 (also when the whole row is masked)
                        if j2 < 2:
                            pv = ll.exp2(s[nt3 * 4 + j2] - mn0)
                        else:
                            pv = ll.exp2(s[nt3 * 4 + j2] - mn1)
                    s[nt3 * 4 + j2] = pv
                    if j2 < 2:
                        rs0 += pv
                    else:
                        rs1 += pv
            l0 = l0 * corr0 + rs0                         # per-lane partial row sums; reduced over the quad once, at the end
            l1 = l1 * corr1 + rs1
            for dt in ll.static_range(16):
                o[dt * 4 + 0] = o[dt * 4 + 0] * corr0
                o[dt * 4 + 1] = o[dt * 4 + 1] * corr0
                o[dt * 4 + 2] = o[dt * 4 + 2] * corr1
                o[dt * 4 + 3] = o[dt * 4 + 3] * corr1
            # ---- O += P V: 4 k-steps of 16 keys, 16 tiles of 8 d's ----
            for kk in ll.static_range(4):
                pa0 = ll.pack_bf16x2(s[(2 * kk) * 4 + 0], s[(2 * kk) * 4 + 1])
                pa1 = ll.pack_bf16x2(s[(2 * kk) * 4 + 2], s[(2 * kk) * 4 + 3])
                pa2 = ll.pack_bf16x2(s[(2 * kk + 1) * 4 + 0], s[(2 * kk + 1) * 4 + 1])
                pa3 = ll.pack_bf16x2(s[(2 * kk + 1) * 4 + 2], s[(2 * kk + 1) * 4 + 3])
                for dt2 in ll.static_range(16):
                    vw = (dt2 * 8 + g) * VS + kk * 8 + tig
                    ll.mma_m16n8k16_bf16(o, dt2 * 4, pa0, pa1, pa2, pa3, Vt[vw], Vt[vw + 4])
            kv0 += BKV
        l0 = l0 + ll.shfl_xor(l0, 1)
        l0 = l0 + ll.shfl_xor(l0, 2)
        l1 = l1 + ll.shfl_xor(l1, 1)
        l1 = l1 + ll.shfl_xor(l1, 2)
        inv0: ll.f32 = 0.0
        inv1: ll.f32 = 0.0
        if l0 > 0.0:
            inv0 = 1.0 / l0
        if l1 > 0.0:
            inv1 = 1.0 / l1
        ow = ll.ptr_cast(out, ll.u32)
        for dt3 in ll.static_range(16):
            col = h * D + dt3 * 8 + tig * 2
            if r0 < S:
                ow[((ll.i64(b) * S + r0) * o_ts + col) // 2] = ll.pack_bf16x2(o[dt3 * 4 + 0] * inv0, o[dt3 * 4 + 1] * inv0)
            if r0 + 8 < S:
                ow[((ll.i64(b) * S + r0 + 8) * o_ts + col) // 2] = ll.pack_bf16x2(o[dt3 * 4 + 2] * inv1, o[dt3 * 4 + 3] * inv1)

    flash_mma.name = f"lk_flash_mma_t{THREADS}"
    flash_mma.bq = BQ
    flash_mma.smem_bytes = (BKV * KS + D * VS) * 4
    return flash_mma


_CACHE = {}


def run_flash_mma(q, k, v, causal: bool = True, sm_scale=None, softcap: float = 0.0, threads: int = 256, interpret: bool = False, out=None):
    """q: [B, S, Hq, 128], k / v: [B, S, Hkv, 128] bf16 -- views into a packed qkv tensor are fine as long as the last two dims are dense
    and q / k / v share the batch stride ``S * token_stride`` -> [B, S, Hq, 128] bf16."""
    import torch
    B, S, Hq, d = q.shape
    Hkv = k.shape[2]
    assert d == D and k.shape == v.shape == (B, S, Hkv, D) and Hq % Hkv == 0
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.stride(3) == 1 and t.stride(2) == D and t.stride(0) == S * t.stride(1)
    assert k.stride(1) == v.stride(1)
    if threads not in _CACHE:
        _CACHE[threads] = make_flash_mma(threads)
    kern = _CACHE[threads]
    out = torch.empty(B, S, Hq, D, dtype=torch.bfloat16, device=q.device) if out is None else out
    grid = (S + kern.bq - 1) // kern.bq * Hq * B
    args = (q, k, v, out, S, Hq, Hkv, q.stride(1), k.stride(1), Hq * D, float(sm_scale if sm_scale is not None else D ** -0.5),
            float(softcap), int(bool(causal)))
    if interpret or not q.is_cuda:
        kern.interpret(grid, *args)
    else:
        kern[grid](*args)
    return out
