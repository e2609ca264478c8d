"""Model glue kernels (csrc/elementwise.cu) with eager fallbacks for the CPU emulation backend.

Reference: flashinfer rmsnorm / rope + HF SiLU in the TP layers (layers/nvidia/tp_attn.py:61-68,165-176,
tp_mlp.py:159), swiglu.py, and the megakernel's norm / activation / rope tasks."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _C

c_void_p, c_ll, c_int, c_float = C.c_void_p, C.c_longlong, C.c_int, C.c_float
_C.register("td_rmsnorm", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_float, c_int, c_void_p])
_C.register("td_silu_mul", c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p])
_C.register("td_silu_mul_bwd", c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p])
_C.register("td_qk_norm_rope_kv", c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_ll, c_float, c_float, c_int, c_void_p])


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, residual: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None):
    """``rmsnorm(x) * w``; with ``residual`` computes ``h = x + residual`` first and returns ``(norm(h), h)``."""
    H = x.shape[-1]
    if not x.is_cuda or x.dtype not in (torch.bfloat16, torch.float16) or H % 8:
        h = x if residual is None else x + residual
        y = (h.float() * torch.rsqrt(h.float().pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype) * weight
        return y if residual is None else (y, h)
    x = x.contiguous()
    rows = x.numel() // H
    out = torch.empty_like(x) if out is None else out
    res_out = torch.empty_like(x) if residual is not None else None
    _C.check(_C.cuda_lib().td_rmsnorm(out.data_ptr(), x.data_ptr(), weight.data_ptr(),
                                      residual.contiguous().data_ptr() if residual is not None else None,
                                      res_out.data_ptr() if res_out is not None else None, rows, H, eps,
                                      int(x.dtype == torch.bfloat16), _s()), "td_rmsnorm")
    return out if residual is None else (out, res_out)


def silu_mul(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``silu(x[..., :I]) * x[..., I:]`` (SwiGLU gate; reference swiglu.py:swiglu_forward)."""
    I = x.shape[-1] // 2
    if not x.is_cuda or x.dtype not in (torch.bfloat16, torch.float16) or I % 8:
        g, u = x[..., :I], x[..., I:]
        return torch.nn.functional.silu(g) * u
    x = x.contiguous()
    M = x.numel() // (2 * I)
    out = torch.empty(x.shape[:-1] + (I,), dtype=x.dtype, device=x.device) if out is None else out
    _C.check(_C.cuda_lib().td_silu_mul(out.data_ptr(), x.data_ptr(), M, I, int(x.dtype == torch.bfloat16), _s()), "td_silu_mul")
    return out


swiglu_forward = silu_mul


def silu_mul_backward(grad_out: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """d/dx of ``silu(x[..., :I]) * x[..., I:]`` in one kernel (reference swiglu.py backward kernels)."""
    I = x.shape[-1] // 2
    if not x.is_cuda or x.dtype not in (torch.bfloat16, torch.float16) or I % 8:
        g, u = x[..., :I].float(), x[..., I:].float()
        sg = torch.sigmoid(g)
        return torch.cat([grad_out.float() * u * (sg + g * sg * (1 - sg)), grad_out.float() * g * sg], dim=-1).to(x.dtype)
    x = x.contiguous()
    gy = grad_out.to(x.dtype).contiguous()
    M = x.numel() // (2 * I)
    dx = torch.empty_like(x)
    _C.check(_C.cuda_lib().td_silu_mul_bwd(dx.data_ptr(), gy.data_ptr(), x.data_ptr(), M, I, int(x.dtype == torch.bfloat16), _s()), "td_silu_mul_bwd")
    return dx


swiglu_backward = silu_mul_backward


def rope_reference(x: torch.Tensor, positions: torch.Tensor, theta: float) -> torch.Tensor:
    """neox-style rotary on ``[T, H, D]`` (fp32 math) -- the eager golden for the fused kernel."""
    D = x.shape[-1]
    inv = theta ** (-torch.arange(0, D, 2, device=x.device, dtype=torch.float32) / D)
    ang = positions.float()[:, None] * inv[None, :]
    cos, sin = ang.cos()[:, None, :], ang.sin()[:, None, :]
    x1, x2 = x.float()[..., :D // 2], x.float()[..., D // 2:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1).to(x.dtype)


def qk_norm_rope_kv(qkv: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, positions: torch.Tensor,
                    batch_idx: Optional[torch.Tensor], Hq: int, Hkv: int, q_norm_w: Optional[torch.Tensor],
                    k_norm_w: Optional[torch.Tensor], eps: float, rope_theta: float) -> torch.Tensor:
    """Fused per-head q/k RMSNorm (optional) + RoPE + KV-cache append.  ``qkv``: [T, (Hq+2Hkv)*128];
    caches: [B, max_len, Hkv, 128]; returns rotated q ``[T, Hq, 128]``."""
    T = qkv.shape[0]
    D = 128
    if not qkv.is_cuda or qkv.dtype not in (torch.bfloat16, torch.float16) or qkv.shape[1] != (Hq + 2 * Hkv) * D:
        D = qkv.shape[1] // (Hq + 2 * Hkv)
        q, k, v = qkv.view(T, Hq + 2 * Hkv, D).split([Hq, Hkv, Hkv], dim=1)
        if q_norm_w is not None:
            q = rmsnorm(q, q_norm_w, eps)
            k = rmsnorm(k, k_norm_w, eps)
        q, k = rope_reference(q, positions, rope_theta), rope_reference(k, positions, rope_theta)
        b = batch_idx.long() if batch_idx is not None else torch.zeros(T, dtype=torch.long, device=qkv.device)
        k_cache[b, positions.long()] = k
        v_cache[b, positions.long()] = v
        return q.contiguous()
    qkv = qkv.contiguous()
    q_out = torch.empty((T, Hq, D), dtype=qkv.dtype, device=qkv.device)
    _C.check(_C.cuda_lib().td_qk_norm_rope_kv(
        qkv.data_ptr(), q_out.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
        q_norm_w.data_ptr() if q_norm_w is not None else None, k_norm_w.data_ptr() if k_norm_w is not None else None,
        positions.data_ptr(), batch_idx.data_ptr() if batch_idx is not None else None, T, Hq, Hkv, k_cache.shape[1],
        eps, rope_theta, int(qkv.dtype == torch.bfloat16), _s()), "td_qk_norm_rope_kv")
    return q_out
