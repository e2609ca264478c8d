"""Gated delta rule (Qwen3-Next linear attention), chunked forward.

Reference: ``chunk_gated_delta_rule_fwd`` (/root/reference/python/triton_dist/kernels/nvidia/gdn.py:926, three Triton
chunk kernels :123,:482,:785).  This is a local (no communication) op and not on any benchmarked path; it is implemented
as the same chunk-parallel algorithm on batched tensor-core matmuls (intra-chunk triangular solve + inter-chunk state
recurrence), fp32 accumulation, so it is exact against the sequential recurrence.

Recurrence per head (S is [Dk, Dv]):   S_t = g_t * S_{t-1} + beta_t * k_t^T (v_t - g_t * k_t S_{t-1}) ;  o_t = q_t S_t
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def gated_delta_rule_recurrent(q, k, v, g, beta, scale: Optional[float] = None, initial_state=None):
    """Sequential reference.  q,k: [B,T,H,Dk]; v: [B,T,H,Dv]; g (log-decay), beta: [B,T,H]."""
    B, T, H, Dk = q.shape
    Dv = v.shape[-1]
    scale = scale if scale is not None else Dk ** -0.5
    S = torch.zeros(B, H, Dk, Dv, dtype=torch.float32, device=q.device) if initial_state is None else initial_state.float().clone()
    o = torch.empty(B, T, H, Dv, dtype=torch.float32, device=q.device)
    qf, kf, vf, gf, bf = q.float() * scale, k.float(), v.float(), g.float().exp(), beta.float()
    for t in range(T):
        S = S * gf[:, t, :, None, None]
        pred = torch.einsum("bhk,bhkv->bhv", kf[:, t], S)
        S = S + torch.einsum("bhk,bhv->bhkv", kf[:, t], (vf[:, t] - pred) * bf[:, t, :, None])
        o[:, t] = torch.einsum("bhk,bhkv->bhv", qf[:, t], S)
    return o.to(q.dtype), S


def fused_recurrent_gated_delta_rule(q, k, v, g, beta, scale: Optional[float] = None, initial_state=None):
    """Token-sequential CUDA kernel (csrc/gdn_kernels.cu): the [Dk, Dv] state lives in registers for the whole call --
    the decode-step / short-prompt companion of the chunked forward.  Returns (o, final_state fp32 [B,H,Dk,Dv])."""
    if not q.is_cuda:
        return gated_delta_rule_recurrent(q, k, v, g, beta, scale, initial_state)
    import ctypes as C
    from .. import _C
    B, T, H, Dk = q.shape
    Dv = v.shape[-1]
    scale = scale if scale is not None else Dk ** -0.5
    state = (torch.zeros(B, H, Dk, Dv, dtype=torch.float32, device=q.device) if initial_state is None
             else initial_state.float().clone().contiguous())
    o = torch.empty(B, T, H, Dv, dtype=q.dtype, device=q.device)
    dt = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[q.dtype]
    fn = _C.cuda_lib().td_gdn_recurrent
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_float, C.c_int, C.c_void_p]
    _C.check(fn(q.contiguous().data_ptr(), k.contiguous().data_ptr(), v.contiguous().data_ptr(), g.float().contiguous().data_ptr(),
                beta.float().contiguous().data_ptr(), state.data_ptr(), o.data_ptr(), B, T, H, Dk, Dv, float(scale), dt,
                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gdn_recurrent")
    return o, state


def chunk_gated_delta_rule_fwd(q, k, v, g, beta, scale: Optional[float] = None, initial_state=None, output_final_state: bool = True,
                               chunk_size: int = 64) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Chunk-parallel forward (WY representation): within a chunk the delta-rule updates are resolved with one
    unit-lower-triangular solve, chunks are chained through the [Dk, Dv] state."""
    B, T, H, Dk = q.shape
    Dv = v.shape[-1]
    if q.is_cuda and T <= 16 and Dk in (64, 128, 256):          # decode steps: state-in-registers kernel
        o, S = fused_recurrent_gated_delta_rule(q, k, v, g, beta, scale, initial_state)
        return o, (S if output_final_state else None)
    import os
    if (q.is_cuda and os.environ.get("TD_GDN_CHUNK_KERNEL", "0") == "1" and Dv <= Dk + 1 and (Dv % 32 == 0 or Dv <= 32)
            and q.dtype in (torch.bfloat16, torch.float16, torch.float32)):
        # prefill on our own kernels: the two DSL kernels of lk/kernels/gdn_chunk.py (chunk preparation in parallel, then the
        # state scan); opt-in until they have run on hardware -- the CPU interpreter reproduces the recurrence to 1e-7
        from ..lk.kernels.gdn_chunk import chunk_gated_delta_rule_lk
        o, S = chunk_gated_delta_rule_lk(q, k, v, g, beta, scale, initial_state, chunk_size)
        return o, (S if output_final_state else None)
    scale = scale if scale is not None else Dk ** -0.5
    C = chunk_size
    pad = (C - T % C) % C
    if pad:
        z = lambda x: torch.nn.functional.pad(x, (0, 0) * (x.dim() - 2) + (0, pad)) if x.dim() == 3 else torch.nn.functional.pad(x, (0, 0, 0, 0, 0, pad))
        q, k, v = z(q), z(k), z(v)
        g = torch.nn.functional.pad(g, (0, 0, 0, pad))
        beta = torch.nn.functional.pad(beta, (0, 0, 0, pad))
    Tp = T + pad
    n = Tp // C
    f = lambda x: x.float().view(B, n, C, H, -1).permute(0, 3, 1, 2, 4)          # [B,H,n,C,D]
    qc, kc, vc = f(q) * scale, f(k), f(v)
    gc = g.float().view(B, n, C, H).permute(0, 3, 1, 2)                          # [B,H,n,C] log decay
    bc = beta.float().view(B, n, C, H).permute(0, 3, 1, 2)
    gcum = gc.cumsum(-1)                                                          # decay from chunk start to t (inclusive)
    # decay between positions inside a chunk: D[i,j] = exp(gcum_i - gcum_j) for j <= i
    diff = gcum[..., :, None] - gcum[..., None, :]
    tril = torch.tril(torch.ones(C, C, device=q.device, dtype=torch.bool))
    Dm = torch.where(tril, diff.exp(), torch.zeros_like(diff))
    kb = kc * bc[..., None]
    # A = strictly-lower (beta_i k_i . k_j decay_ij);  (I + A) u = beta * (v - decayed k S0)  ->  solve once per chunk
    A = torch.einsum("bhnik,bhnjk->bhnij", kb, kc) * Dm
    A = A * torch.tril(torch.ones(C, C, device=q.device), -1)
    eye = torch.eye(C, device=q.device).expand_as(A)
    Tinv = torch.linalg.solve_triangular(eye + A, eye, upper=False)              # [B,H,n,C,C]
    w = Tinv @ (kb * gcum.exp()[..., None])                                       # multiplies the incoming state
    u = Tinv @ (vc * bc[..., None])
    S = torch.zeros(B, H, Dk, Dv, dtype=torch.float32, device=q.device) if initial_state is None else initial_state.float().clone()
    out = torch.empty(B, H, n, C, Dv, dtype=torch.float32, device=q.device)
    qk = torch.einsum("bhnik,bhnjk->bhnij", qc, kc) * Dm
    for c in range(n):
        delta = u[:, :, c] - w[:, :, c] @ S                                       # effective "new values" of this chunk
        out[:, :, c] = (qc[:, :, c] * gcum[:, :, c].exp()[..., None]) @ S + qk[:, :, c] @ delta
        decay_end = (gcum[:, :, c, -1:, None] - gcum[:, :, c, :, None]).exp()     # from position j to the chunk end
        S = S * gcum[:, :, c, -1].exp()[..., None, None] + (kc[:, :, c] * decay_end).transpose(-1, -2) @ delta
    o = out.permute(0, 2, 3, 1, 4).reshape(B, Tp, H, Dv)[:, :T].to(v.dtype)
    return o, (S if output_final_state else None)


# ---- variable-length bookkeeping of the chunked forward (reference: kernels/nvidia/gdn.py prepare_lens / prepare_chunk_indices / prepare_chunk_offsets) --
def prepare_lens(cu_seqlens: torch.Tensor) -> torch.Tensor:
    """Sequence lengths from cumulative lengths [B + 1]."""
    return cu_seqlens[1:] - cu_seqlens[:-1]


def prepare_chunk_indices(cu_seqlens: torch.Tensor, chunk_size: int) -> torch.Tensor:
    """[total_chunks, 2] rows ``(sequence, chunk inside the sequence)``: the work list of a kernel that takes one chunk per CTA over a
    packed batch.  Device-side (repeat_interleave + arange), no host read of the lengths besides the output size."""
    n = (prepare_lens(cu_seqlens) + chunk_size - 1) // chunk_size                       # chunks per sequence
    seq = torch.repeat_interleave(torch.arange(n.numel(), device=cu_seqlens.device), n)
    first = torch.cumsum(n, 0) - n
    within = torch.arange(seq.numel(), device=cu_seqlens.device) - first[seq]
    return torch.stack([seq, within], 1).to(cu_seqlens.dtype)


def prepare_chunk_offsets(cu_seqlens: torch.Tensor, chunk_size: int) -> torch.Tensor:
    """[B + 1] cumulative chunk counts: chunks of sequence b are ``offsets[b] .. offsets[b + 1]`` of the packed chunk list."""
    n = (prepare_lens(cu_seqlens) + chunk_size - 1) // chunk_size
    return torch.cat([n.new_zeros(1), torch.cumsum(n, 0)]).to(cu_seqlens.dtype)

