"""GQA decode attention (split-KV + combine) in the DSL, D = 128.

Reference: kernels/nvidia/flash_decode.py ``kernel_gqa_fwd_batch_decode_split_kv`` (:130) / ``..._combine_kv`` (:308) -- one query token per
sequence attends to a long KV cache; the keys are split so that a handful of (batch, kv head) pairs still fills the GPU, every split
produces an (m, l, o) partial, a second kernel merges them; ``kernel_inter_rank_gqa_fwd_batch_decode_combine_kv`` (:482) merges partials of
KV-sharded ranks the same way.  The product's CUDA kernels are csrc/attention.cu (flash-decode v2) and the megakernel's ATTN task; this
is that task's algorithm as two DSL kernels:

* ``decode_split``: a CTA per (batch, kv head, split); a warp per key (lane = 4 of the 128 dims, 8-byte loads), scores by warp-shuffle
  reduction for each of the G query heads of the group, per-warp online softmax, warps merged through shared memory; writes the partial
  ``[m, l, o(128)]`` per query head -- or, with one split, the normalised output directly;
* ``decode_combine``: a CTA per (batch, kv head): log-sum-exp merge of the splits (also used to merge the partials gathered from other
  ranks: the merge is associative, so local splits and remote shards are the same thing).

Both run in the CPU interpreter against the softmax reference (tests/test_lk_cpu.py); soft-cap and paged caches are left to the CUDA kernels.
"""
from triton_dist import lk
from triton_dist.lk import ll

THREADS = 256
NW = THREADS // 32
GMAX = 8                       # query heads per kv head
PART = 130                     # floats per partial: m, l, o[128]


@lk.kernel(block=THREADS)
def decode_split(q: ll.ptr[ll.bf16], kc: ll.ptr[ll.bf16], vc: ll.ptr[ll.bf16], kv_lens: ll.ptr[ll.i32], part: ll.ptr[ll.f32],
                 out: ll.ptr[ll.bf16], Hq: ll.i32, Hkv: ll.i32, max_len: ll.i32, n_splits: ll.i32, scale: ll.f32):
    """q: [B, Hq, 128]; kc / vc: [B, max_len, Hkv, 128]; part: [B, Hkv, n_splits, GMAX, 130]; out: [B, Hq, 128] (written when n_splits == 1)."""
    sm_m = ll.shared([NW * GMAX], ll.f32)
    sm_l = ll.shared([NW * GMAX], ll.f32)
    sm_o = ll.shared([NW * GMAX * 128], ll.f32)
    tid = ll.threadIdx.x
    warp = tid // 32
    lane = tid % 32
    split = ll.blockIdx.x % n_splits
    kvh = (ll.blockIdx.x // n_splits) % Hkv
    b = ll.blockIdx.x // (n_splits * Hkv)
    G = Hq // Hkv
    length = kv_lens[b]
    per = (length + n_splits - 1) // n_splits
    j0 = split * per
    j1 = min(length, j0 + per)
    qf = ll.local([GMAX * 4], ll.f32)
    m = ll.local([GMAX], ll.f32)
    l = ll.local([GMAX], ll.f32)
    o = ll.local([GMAX * 4], ll.f32)
    for g in ll.static_range(GMAX):
        m[g] = -1.0e30
        l[g] = 0.0
        for e in ll.static_range(4):
            o[g * 4 + e] = 0.0
            qf[g * 4 + e] = 0.0
        if g < G:
            for e2 in ll.static_range(4):
                qf[g * 4 + e2] = ll.f32(q[(ll.i64(b) * Hq + kvh * G + g) * 128 + lane * 4 + e2]) * scale
    for j in range(j0 + warp, j1, NW):
        row = ((ll.i64(b) * max_len + j) * Hkv + kvh) * 128 + lane * 4
        k0 = ll.f32(kc[row + 0])
        k1 = ll.f32(kc[row + 1])
        k2 = ll.f32(kc[row + 2])
        k3 = ll.f32(kc[row + 3])
        v0 = ll.f32(vc[row + 0])
        v1 = ll.f32(vc[row + 1])
        v2 = ll.f32(vc[row + 2])
        v3 = ll.f32(vc[row + 3])
        for g2 in ll.static_range(GMAX):
            if g2 < G:
                s = qf[g2 * 4] * k0 + qf[g2 * 4 + 1] * k1 + qf[g2 * 4 + 2] * k2 + qf[g2 * 4 + 3] * k3
                for off in ll.static_range(5):
                    s += ll.shfl_xor(s, 1 << off)
                mn = max(m[g2], s)
                corr = ll.exp(m[g2] - mn)
                p = ll.exp(s - mn)
                l[g2] = l[g2] * corr + p
                o[g2 * 4 + 0] = o[g2 * 4 + 0] * corr + p * v0
                o[g2 * 4 + 1] = o[g2 * 4 + 1] * corr + p * v1
                o[g2 * 4 + 2] = o[g2 * 4 + 2] * corr + p * v2
                o[g2 * 4 + 3] = o[g2 * 4 + 3] * corr + p * v3
                m[g2] = mn
    for g3 in ll.static_range(GMAX):
        if g3 < G:
            if lane == 0:
                sm_m[warp * GMAX + g3] = m[g3]
                sm_l[warp * GMAX + g3] = l[g3]
            for e3 in ll.static_range(4):
                sm_o[(warp * GMAX + g3) * 128 + lane * 4 + e3] = o[g3 * 4 + e3]
    ll.syncthreads()
    for idx in range(tid, G * 128, THREADS):
        gg = idx // 128
        d = idx % 128
        mm: ll.f32 = -1.0e30
        for w in range(NW):
            mm = max(mm, sm_m[w * GMAX + gg])
        lsum: ll.f32 = 0.0
        osum: ll.f32 = 0.0
        for w2 in range(NW):
            c = ll.exp(sm_m[w2 * GMAX + gg] - mm)
            lsum += sm_l[w2 * GMAX + gg] * c
            osum += sm_o[(w2 * GMAX + gg) * 128 + d] * c
        if n_splits > 1:
            base = (((ll.i64(b) * Hkv + kvh) * n_splits + split) * GMAX + gg) * PART
            if d == 0:
                part[base] = mm
                part[base + 1] = lsum
            part[base + 2 + d] = osum
        else:
            inv: ll.f32 = 0.0
            if lsum > 0.0:
                inv = 1.0 / lsum
            out[(ll.i64(b) * Hq + kvh * G + gg) * 128 + d] = osum * inv


@lk.kernel(block=THREADS)
def decode_combine(part: ll.ptr[ll.f32], out: ll.ptr[ll.bf16], lse: ll.ptr[ll.f32], Hq: ll.i32, Hkv: ll.i32, n_parts: ll.i32):
    """part: [B, Hkv, n_parts, GMAX, 130] -> out [B, Hq, 128] (bf16) and lse [B, Hq] = log-sum-exp of the scores (for a further merge)."""
    tid = ll.threadIdx.x
    kvh = ll.blockIdx.x % Hkv
    b = ll.blockIdx.x // Hkv
    G = Hq // Hkv
    for idx in range(tid, G * 128, THREADS):
        gg = idx // 128
        d = idx % 128
        mm: ll.f32 = -1.0e30
        for s in range(n_parts):
            mm = max(mm, part[(((ll.i64(b) * Hkv + kvh) * n_parts + s) * GMAX + gg) * PART])
        lsum: ll.f32 = 0.0
        osum: ll.f32 = 0.0
        for s2 in range(n_parts):
            base = (((ll.i64(b) * Hkv + kvh) * n_parts + s2) * GMAX + gg) * PART
            c = ll.exp(part[base] - mm)
            lsum += part[base + 1] * c
            osum += part[base + 2 + d] * c
        inv: ll.f32 = 0.0
        if lsum > 0.0:
            inv = 1.0 / lsum
        out[(ll.i64(b) * Hq + kvh * G + gg) * 128 + d] = osum * inv
        if d == 0:
            lse[ll.i64(b) * Hq + kvh * G + gg] = mm + ll.log(max(lsum, 1.0e-38))


def gqa_decode_lk(q, k_cache, v_cache, kv_lens, n_splits: int = 4, sm_scale=None, interpret: bool = False, return_lse: bool = False):
    """q: [B, Hq, 128] bf16, k / v cache: [B, max_len, Hkv, 128] bf16, kv_lens: int32 [B] -> [B, Hq, 128] bf16 (and lse [B, Hq])."""
    import torch
    B, Hq, D = q.shape
    max_len, Hkv = k_cache.shape[1], k_cache.shape[2]
    assert D == 128 and Hq % Hkv == 0 and Hq // Hkv <= GMAX and q.dtype == torch.bfloat16 and k_cache.dtype == torch.bfloat16
    scale = float(sm_scale if sm_scale is not None else D ** -0.5)
    ns = max(1, n_splits)
    if return_lse and ns < 2:
        ns = 2                                   # the log-sum-exp comes out of the combine kernel: go through the partial path
    out = torch.empty(B, Hq, D, dtype=torch.bfloat16, device=q.device)
    lse = torch.zeros(B, Hq, dtype=torch.float32, device=q.device)
    part = torch.zeros(B * Hkv * ns * GMAX * PART, dtype=torch.float32, device=q.device)
    run = (lambda k, grid, *a: k.interpret(grid, *a)) if (interpret or not q.is_cuda) else (lambda k, grid, *a: k[grid](*a))
    run(decode_split, B * Hkv * ns, q.contiguous(), k_cache.contiguous(), v_cache.contiguous(), kv_lens.to(torch.int32), part, out, Hq, Hkv,
        max_len, ns, scale)
    if ns > 1:
        run(decode_combine, B * Hkv, part, out, lse, Hq, Hkv, ns)
    return (out, lse) if return_lse else out


class LkSpDecode:
    """KV-sharded (sequence-parallel) decode entirely on DSL kernels: every rank attends to ITS shard of the KV cache (``decode_split`` +
    ``decode_combine`` -> output and log-sum-exp), the per-head partials ``(m = lse, l = 1, o)`` are exchanged with the low-latency
    flag-in-data all-gather (``allgather_ll``), and ``decode_combine`` merges the W partials.  Reference: layers/nvidia/
    sp_flash_decode_layer.py ``SpGQAFlashDecodeAttention`` (local split-KV decode -> LL all-gather of (O, LSE) -> inter-rank combine)."""

    def __init__(self, batch: int, num_q_heads: int, num_kv_heads: int, n_splits: int = 2):
        import triton_dist.utils as U
        from .allgather_ll import LkLLAllGather
        self.B, self.Hq, self.Hkv, self.n_splits = batch, num_q_heads, num_kv_heads, n_splits
        self.W = U.world_size()
        self.ag = LkLLAllGather(batch * num_kv_heads * GMAX * PART * 4)

    def __call__(self, q, k_shard, v_shard, local_kv_lens, sm_scale=None):
        """q: [B, Hq, 128] (replicated); k / v shard: [B, max_len_local, Hkv, 128]; local_kv_lens: int32 [B] keys of every sequence that
        live on THIS rank (may be 0) -> [B, Hq, 128]."""
        import torch
        B, Hq, Hkv, W = self.B, self.Hq, self.Hkv, self.W
        G = Hq // Hkv
        o, lse = gqa_decode_lk(q, k_shard, v_shard, local_kv_lens, n_splits=self.n_splits, sm_scale=sm_scale, return_lse=True)
        empty = (local_kv_lens.to(q.device) <= 0)[:, None].expand(B, Hq)                 # a rank without keys contributes weight exp(-inf) = 0
        mine = torch.zeros(B, Hkv, GMAX, PART, dtype=torch.float32, device=q.device)
        mine[:, :, :G, 0] = torch.where(empty, torch.full_like(lse, -1.0e30), lse).view(B, Hkv, G)
        mine[:, :, :G, 1] = torch.where(empty, torch.zeros_like(lse), torch.ones_like(lse)).view(B, Hkv, G)
        mine[:, :, :G, 2:] = o.float().view(B, Hkv, G, 128)
        allp = self.ag(mine.view(-1))                                                    # [W, B * Hkv * GMAX * PART]
        part = allp.view(W, B, Hkv, GMAX, PART).permute(1, 2, 0, 3, 4).contiguous().view(-1)
        out = torch.empty(B, Hq, 128, dtype=torch.bfloat16, device=q.device)
        lse_all = torch.empty(B, Hq, dtype=torch.float32, device=q.device)
        if q.is_cuda:
            decode_combine[B * Hkv](part, out, lse_all, Hq, Hkv, W)
        else:
            decode_combine.interpret(B * Hkv, part, out, lse_all, Hq, Hkv, W)
        return out

    def finalize(self):
        self.ag.finalize()

