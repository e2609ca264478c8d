"""Chunked gated-delta-rule forward (Qwen3-Next style linear attention prefill) as two DSL kernels.

    S_t = exp(g_t) S_{t-1} + beta_t k_t^T (v_t - k_t exp(g_t) S_{t-1}),        o_t = scale * q_t S_t

The reference runs six Triton kernels (kernels/nvidia/gdn.py: chunk-local cumsum, scaled k k^T, solve_tril, recompute w / u, chunk_fwd_h,
chunk_fwd_o).  Here the same WY-form algorithm is two kernels written in the ``triton_dist.lk`` DSL:

* ``prepare``  (one CTA per chunk, all chunks in parallel): the in-chunk decay prefix, A = strict_lower(beta_i k_i.k_j exp(G_i - G_j)),
  T = (I + A)^-1 by forward substitution in shared memory, w = T (beta k exp(G)), u = T (beta v), the masked q k^T block, and the
  pre-scaled q exp(G) / k exp(G_end - G) rows the scan needs;
* ``scan``     (one CTA per (batch, head, slice of DS value columns), chunks in order): delta = u - w S, o = qg S + qk delta,
  S <- exp(G_end) S + kd^T delta with the [Dk, DS] state resident in shared memory for the whole sequence.

Both are plain SIMT fp32 kernels (the op is a small fraction of a layer's FLOPs and has a sequential dependency per chunk); being DSL
kernels they also run in the CPU interpreter, which is how ``tests/test_lk_cpu.py`` checks them against the token-by-token recurrence.
``ops/gdn.py`` uses them for CUDA prefill when ``TD_GDN_CHUNK_KERNEL=1`` (opt-in until they have run on hardware).
"""
# no ``from __future__ import annotations`` here: the kernel signatures use the factory parameter ``dtype``, which only exists as an
# evaluated annotation object (it is not referenced in the kernel bodies, so it is not a closure variable either)
from triton_dist import lk
from triton_dist.lk import ll

THREADS = 256


def make_gdn_chunk_kernels(C: int = 64, DK: int = 128, DV: int = 128, DS: int = 32, dtype=ll.bf16):
    """Kernel factory: chunk length, head dims, value-column slice per scan CTA and the q / k / v / o element type are compile-time."""
    assert THREADS % DS == 0 and DV % DS == 0 and C <= THREADS
    RPT = THREADS // DS                        # rows of a [*, DS] tile covered per pass by the (row = tid / DS, col = tid % DS) mapping
    KP, CP = DK + 1, C + 1                     # padded shared-memory row lengths (conflict-free column walks)

    @lk.kernel(block=THREADS)
    def prepare(q: ll.ptr[dtype], k: ll.ptr[dtype], v: ll.ptr[dtype], g: ll.ptr[ll.f32], beta: ll.ptr[ll.f32],
                w: ll.ptr[ll.f32], u: ll.ptr[ll.f32], qk: ll.ptr[ll.f32], qg: ll.ptr[ll.f32], kd: ll.ptr[ll.f32], gend: ll.ptr[ll.f32],
                T: ll.i32, H: ll.i32, n_chunks: ll.i32, scale: ll.f32):
        ks = ll.dyn_shared([C, KP], ll.f32)            # k rows of the chunk
        xs = ll.dyn_shared([C, KP], ll.f32)            # q rows, later beta * k * exp(G)
        A = ll.dyn_shared([C, CP], ll.f32)
        Ti = ll.dyn_shared([C, CP], ll.f32)            # (I + A)^-1
        G = ll.dyn_shared([C], ll.f32)                 # inclusive prefix of the log-decays inside the chunk
        bt = ll.dyn_shared([C], ll.f32)
        tid = ll.threadIdx.x
        chunk = ll.blockIdx.x
        h = ll.blockIdx.y
        b = ll.blockIdx.z
        t0 = chunk * C
        bh = ll.i64(b) * H + h
        cbase = (bh * n_chunks + chunk) * C            # first row of this chunk in the [B*H, n, C, *] intermediates

        # ---- load: rows past the end of the sequence are zeros with g = 0, beta = 0 (they change nothing) ----
        for i in range(tid, C * DK, THREADS):
            r = i // DK
            d = i % DK
            tok = t0 + r
            src = ((ll.i64(b) * T + tok) * H + h) * DK + d
            kv: ll.f32 = 0.0
            qv: ll.f32 = 0.0
            if tok < T:
                kv = k[src]
                qv = q[src]
            ks[r, d] = kv
            xs[r, d] = qv * scale
        if tid < C:
            tok1 = t0 + tid
            gv: ll.f32 = 0.0
            bv: ll.f32 = 0.0
            if tok1 < T:
                gv = g[(ll.i64(b) * T + tok1) * H + h]
                bv = beta[(ll.i64(b) * T + tok1) * H + h]
            G[tid] = gv
            bt[tid] = bv
        ll.syncthreads()
        if tid == 0:                                    # 64-element inclusive scan: one thread, C adds
            run: ll.f32 = 0.0
            for i in range(C):
                run += G[i]
                G[i] = run
            gend[bh * n_chunks + chunk] = ll.exp(run)
        ll.syncthreads()

        # ---- A (strictly lower) and the masked q k^T block ----
        for e in range(tid, C * C, THREADS):
            i = e // C
            j = e % C
            dk_: ll.f32 = 0.0
            dq_: ll.f32 = 0.0
            if j <= i:
                for d in range(DK):
                    kj = ks[j, d]
                    dk_ += ks[i, d] * kj
                    dq_ += xs[i, d] * kj
            dec = ll.exp(G[i] - G[j]) if j <= i else 0.0
            A[i, j] = bt[i] * dk_ * dec if j < i else 0.0
            qk[(cbase + i) * C + j] = dq_ * dec
        # q exp(G) and k exp(G_end - G) rows for the scan (q rows are not needed in shared memory after this)
        for i in range(tid, C * DK, THREADS):
            r = i // DK
            d = i % DK
            qg[(cbase + r) * DK + d] = xs[r, d] * ll.exp(G[r])
            kd[(cbase + r) * DK + d] = ks[r, d] * ll.exp(G[C - 1] - G[r])
        ll.syncthreads()

        # ---- Ti = (I + A)^-1: row i of the inverse needs rows < i (forward substitution), column per thread ----
        for i in range(C):
            if tid < C:
                acc: ll.f32 = 1.0 if tid == i else 0.0
                for j in range(i):
                    acc -= A[i, j] * Ti[j, tid]
                Ti[i, tid] = acc
            ll.syncthreads()

        # ---- w = Ti (beta k exp(G)),  u = Ti (beta v) ----
        for i in range(tid, C * DK, THREADS):
            r = i // DK
            d = i % DK
            xs[r, d] = ks[r, d] * bt[r] * ll.exp(G[r])
        ll.syncthreads()
        for e in range(tid, C * DK, THREADS):
            i = e // DK
            d = e % DK
            s: ll.f32 = 0.0
            for j in range(i + 1):
                s += Ti[i, j] * xs[j, d]
            w[(cbase + i) * DK + d] = s
        ll.syncthreads()
        for i in range(tid, C * DV, THREADS):           # beta * v into the (now free) k buffer; DV <= DK + 1 columns fit
            r = i // DV
            d = i % DV
            tok2 = t0 + r
            vv: ll.f32 = 0.0
            if tok2 < T:
                vv = v[((ll.i64(b) * T + tok2) * H + h) * DV + d]
            ks[r, d] = vv * bt[r]
        ll.syncthreads()
        for e in range(tid, C * DV, THREADS):
            i = e // DV
            d = e % DV
            s2: ll.f32 = 0.0
            for j in range(i + 1):
                s2 += Ti[i, j] * ks[j, d]
            u[(cbase + i) * DV + d] = s2

    @lk.kernel(block=THREADS)
    def scan(w: ll.ptr[ll.f32], u: ll.ptr[ll.f32], qk: ll.ptr[ll.f32], qg: ll.ptr[ll.f32], kd: ll.ptr[ll.f32], gend: ll.ptr[ll.f32],
             state: ll.ptr[ll.f32], o: ll.ptr[dtype], T: ll.i32, H: ll.i32, n_chunks: ll.i32):
        S = ll.dyn_shared([DK, DS + 1], ll.f32)         # running state, this CTA's DS value columns
        dl = ll.dyn_shared([C, DS + 1], ll.f32)         # delta of the current chunk
        tid = ll.threadIdx.x
        col = tid % DS
        row0 = tid // DS
        vs = ll.blockIdx.x                              # value-column slice
        h = ll.blockIdx.y
        b = ll.blockIdx.z
        bh = ll.i64(b) * H + h
        c0 = vs * DS
        st = state + bh * DK * DV                       # fp32 [DK, DV], read at the start, written at the end
        for d in range(row0, DK, RPT):
            S[d, col] = st[ll.i64(d) * DV + c0 + col]
        ll.syncthreads()
        for chunk in range(n_chunks):
            cbase = (bh * n_chunks + chunk) * C
            # delta = u - w S
            for i in range(row0, C, RPT):
                acc = u[(cbase + i) * DV + c0 + col]
                wrow = w + (cbase + i) * DK
                for d in range(DK):
                    acc -= wrow[d] * S[d, col]
                dl[i, col] = acc
            ll.syncthreads()
            # o = qg S + qk delta   (qk is lower triangular: j <= i)
            for i in range(row0, C, RPT):
                tok = chunk * C + i
                if tok < T:
                    out: ll.f32 = 0.0
                    qrow = qg + (cbase + i) * DK
                    for d in range(DK):
                        out += qrow[d] * S[d, col]
                    prow = qk + (cbase + i) * C
                    for j in range(i + 1):
                        out += prow[j] * dl[j, col]
                    o[((ll.i64(b) * T + tok) * H + h) * DV + c0 + col] = out
            ll.syncthreads()
            # S <- exp(G_end) S + kd^T delta
            ge = gend[bh * n_chunks + chunk]
            for d in range(row0, DK, RPT):
                acc2 = S[d, col] * ge
                for j in range(C):
                    acc2 += kd[(cbase + j) * DK + d] * dl[j, col]
                S[d, col] = acc2
            ll.syncthreads()
        for d in range(row0, DK, RPT):
            st[ll.i64(d) * DV + c0 + col] = S[d, col]

    prepare.name = f"lk_gdn_prepare_c{C}_k{DK}_v{DV}_{dtype.name}"
    scan.name = f"lk_gdn_scan_c{C}_k{DK}_v{DV}_s{DS}_{dtype.name}"
    return prepare, scan


_CACHE = {}


def get_kernels(C, DK, DV, DS, dtype):
    key = (C, DK, DV, DS, dtype.name)
    if key not in _CACHE:
        _CACHE[key] = make_gdn_chunk_kernels(C, DK, DV, DS, dtype)
    return _CACHE[key]


def chunk_gated_delta_rule_lk(q, k, v, g, beta, scale=None, initial_state=None, chunk_size: int = 64, interpret: bool = False):
    """q, k: [B, T, H, Dk]; v: [B, T, H, Dv]; g (log decay), beta: [B, T, H].  Returns (o [B, T, H, Dv], final state fp32 [B, H, Dk, Dv]).
    ``interpret=True`` runs both kernels in the CPU interpreter (tiny shapes only)."""
    import torch
    B, T, H, DK = q.shape
    DV = v.shape[-1]
    C = chunk_size
    n = (T + C - 1) // C
    scale = float(scale if scale is not None else DK ** -0.5)
    DS = 32 if DV % 32 == 0 else DV
    dt = {torch.bfloat16: ll.bf16, torch.float16: ll.f16, torch.float32: ll.f32}[q.dtype]
    assert DV <= DK + 1, "the prepare kernel reuses the k tile for beta * v"
    prepare, scan = get_kernels(C, DK, DV, DS, dt)
    dev = q.device
    f32 = dict(dtype=torch.float32, device=dev)
    w = torch.empty(B * H * n * C * DK, **f32)
    qg, kd = torch.empty_like(w), torch.empty_like(w)
    u = torch.empty(B * H * n * C * DV, **f32)
    qk = torch.empty(B * H * n * C * C, **f32)
    gend = torch.empty(B * H * n, **f32)
    state = torch.zeros(B, H, DK, DV, **f32) if initial_state is None else initial_state.to(torch.float32).clone().contiguous()
    o = torch.empty(B, T, H, DV, dtype=q.dtype, device=dev)
    qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
    gc, bc = g.to(torch.float32).contiguous(), beta.to(torch.float32).contiguous()
    a1 = (qc, kc, vc, gc, bc, w, u, qk, qg, kd, gend, T, H, n, scale)
    a2 = (w, u, qk, qg, kd, gend, state, o, T, H, n)
    if interpret:
        prepare.interpret((n, H, B), *a1)
        scan.interpret((DV // DS, H, B), *a2)
    else:
        prepare[(n, H, B)](*a1)
        scan[(DV // DS, H, B)](*a2)
    return o, state
