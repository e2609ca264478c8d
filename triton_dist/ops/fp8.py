"""Block-scaled fp8 (MXFP8: e4m3 data + one UE8M0 power-of-two scale per 32 K-elements) for the tcgen05 GEMM family.

``tcgen05.mma.kind::mxf8f6f4.block_scale`` consumes the scales from TMEM, so quantisation writes them directly in the
512-byte tile order the kernel copies with ``tcgen05.cp`` (csrc/quant_kernels.cu, csrc/gemm_sm100.cuh kFP8).
The reference only has per-tensor fp8 through ``tl.dot`` (test_gemm_rs.py:130-145); BASELINE config #3 asks for
block-scaled fp8 on the GEMM-RS path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _C
from .. import utils as U
from .gemm import GemmConfig, fill_common

c_void_p, c_int, c_ll = C.c_void_p, C.c_int, C.c_longlong
_C.register("td_quant_mxfp8", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p])


@dataclass
class MXFP8Tensor:
    q: torch.Tensor        # [M, K] float8_e4m3fn
    sf: torch.Tensor       # uint8 [ceil(M/128), K/128, 512]  (tiled UE8M0 exponents)
    shape: tuple

    @property
    def chunks(self) -> int:
        return self.sf.shape[0] * self.sf.shape[1]


def _sf_index(M: int, K: int, device):
    rows = torch.arange(M, device=device)[:, None]
    kb = torch.arange(K // 32, device=device)[None, :]
    chunk = (rows // 128) * (K // 128) + kb // 4
    return chunk * 512 + (rows % 32) * 16 + ((rows % 128) // 32) * 4 + kb % 4


def quantize_mxfp8(x: torch.Tensor) -> MXFP8Tensor:
    """Row-wise MXFP8 quantisation of a K-major ``[M, K]`` matrix (K % 128 == 0)."""
    M, K = x.shape
    assert K % 128 == 0 and x.stride(1) == 1
    mt, kt = (M + 127) // 128, K // 128
    sf = torch.zeros((mt, kt, 512), dtype=torch.uint8, device=x.device)
    if x.is_cuda and x.dtype in (torch.bfloat16, torch.float16):
        q = torch.empty((M, K), dtype=torch.float8_e4m3fn, device=x.device)
        _C.check(_C.cuda_lib().td_quant_mxfp8(x.data_ptr(), q.data_ptr(), sf.data_ptr(), M, K, x.stride(0),
                                              int(x.dtype == torch.bfloat16), c_void_p(torch.cuda.current_stream().cuda_stream)),
                 "td_quant_mxfp8")
        return MXFP8Tensor(q, sf, (M, K))
    xf = x.float().view(M, K // 32, 32)
    amax = xf.abs().amax(-1)
    e = torch.where(amax > 0, torch.ceil(torch.log2(amax / 448.0)), torch.full_like(amax, -127)).clamp(-127, 127)
    q = (xf * torch.exp2(-e)[..., None]).view(M, K).to(torch.float8_e4m3fn)
    sf.view(-1)[_sf_index(M, K, x.device).reshape(-1)] = (e + 127).to(torch.uint8).reshape(-1)
    return MXFP8Tensor(q, sf, (M, K))


def dequantize_mxfp8(t: MXFP8Tensor) -> torch.Tensor:
    """fp32 reconstruction (the golden the block-scaled GEMM is compared against)."""
    M, K = t.shape
    e = t.sf.view(-1)[_sf_index(M, K, t.q.device).reshape(-1)].view(M, K // 32).float() - 127.0
    return (t.q.float().view(M, K // 32, 32) * torch.exp2(e)[..., None]).view(M, K)


def gemm_mxfp8(a: MXFP8Tensor, b: MXFP8Tensor, out: Optional[torch.Tensor] = None, config: Optional[GemmConfig] = None) -> torch.Tensor:
    """``out[M, N] (bf16) = dequant(a)[M, K] @ dequant(b)[N, K].T`` on the block-scaled tensor-core path."""
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    if not a.q.is_cuda:
        res = (dequantize_mxfp8(a) @ dequantize_mxfp8(b).t()).to(torch.bfloat16)
        if out is not None:
            out.copy_(res)
            return out
        return res
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.q.device) if out is None else out
    cfg = config or GemmConfig(bn=128, cta_group=2 if M >= 256 else 1, group_m=8, use_tma_store=True)
    args = _C.GemmArgs()
    args.mode = 0
    fill_common(args, M, a.q.data_ptr(), a.q.stride(0), b.q, out.data_ptr(), M, out.stride(0), M, N, K, cfg, True)
    fill_fp8(args, a, b)
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemm_launch(mxfp8)")
    return out


def fill_fp8(args: _C.GemmArgs, a: MXFP8Tensor, b: MXFP8Tensor):
    args.is_bf16 = 2
    args.sfa, args.sfb = a.sf.data_ptr(), b.sf.data_ptr()
    args.sfa_chunks, args.sfb_chunks = a.chunks, b.chunks
