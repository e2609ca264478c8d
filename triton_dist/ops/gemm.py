"""Local GEMM on the tcgen05/TMEM/TMA kernel (csrc/gemm_sm100.cuh) -- ``C = A @ B.T``.

Reference counterparts: ``matmul*`` in /root/reference/python/triton_dist/kernels/nvidia/gemm.py:396-875
(Triton tl.dot, per-device config tables with no B200 entry) and little_kernel's gemm_sm100 ladder
(/root/reference/python/little_kernel/benchmark/gemm_sm100/gemm_level9.py: 2-CTA 256x256x64, 4 stages).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _C


@dataclass(frozen=True)
class GemmConfig:
    """Tile configuration of the sm_100a GEMM.  ``bn``: 32/64/128/256, ``cta_group``: 1 or 2 (CTA pair)."""
    bn: int = 256
    cta_group: int = 2
    group_m: int = 8
    use_tma_store: bool = True
    num_sms: int = 0          # 0 = all SMs
    n_comm_ctas: int = 0      # fused ops only

    def key(self):
        return (self.bn, self.cta_group, self.group_m, int(self.use_tma_store), self.num_sms, self.n_comm_ctas)


def default_config(M: int, N: int, K: int, num_sms: int = 148) -> GemmConfig:
    """Shape heuristic (stands in for the reference's per-device config tables, gemm.py:184-393)."""
    if M <= 128:
        # skinny (decode): stream the weights with as many CTAs as possible
        for bn in (32, 64, 128):
            if (N + bn - 1) // bn <= num_sms or bn == 128:
                return GemmConfig(bn=bn, cta_group=1, group_m=1)
    tiles_256 = ((M + 255) // 256) * ((N + 255) // 256)
    if tiles_256 >= num_sms // 2:
        return GemmConfig(bn=256, cta_group=2, group_m=8)
    tiles_128 = ((M + 255) // 256) * ((N + 127) // 128)
    if tiles_128 >= num_sms // 2:
        return GemmConfig(bn=128, cta_group=2, group_m=8)
    return GemmConfig(bn=128, cta_group=1, group_m=8)


def _check_operand(x: torch.Tensor, name: str):
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D and K-major (row-major [rows, K])")
    if x.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError(f"{name}: only bf16/fp16 supported by this entry point, got {x.dtype}")
    if x.data_ptr() % 16 or x.stride(0) % 8:
        raise ValueError(f"{name} must be 16-byte aligned with a leading dimension that is a multiple of 8")


def fill_common(args: _C.GemmArgs, a_rows: int, A_ptr: int, lda: int, B: torch.Tensor, C_ptr: int, c_rows: int, ldc: int,
                M: int, N: int, K: int, cfg: GemmConfig, is_bf16: bool):
    args.is_bf16 = 1 if is_bf16 else 0
    args.bn, args.cta_group, args.group_m = cfg.bn, cfg.cta_group, cfg.group_m
    args.n_comm_ctas, args.use_tma_store, args.num_sms = cfg.n_comm_ctas, int(cfg.use_tma_store), cfg.num_sms
    args.M, args.N, args.K = M, N, K
    args.A, args.a_rows, args.lda, args.a_nbuf, args.a_buf_stride_bytes = A_ptr, a_rows, lda, 1, 0
    args.B, args.ldb = B.data_ptr(), B.stride(0)
    args.C, args.c_rows, args.ldc = C_ptr, c_rows, ldc
    args.world = 1


_SK_SCRATCH = {}
_SK_WS_BYTES = 20 << 20      # >= workers * 128 * cta_group * BN * 4 for every tile shape (74 pairs x 256 x 256 fp32 = 19.4 MB)
_SK_FLAGS = 1024


def attach_splitk(args: _C.GemmArgs, device: torch.device, max_parts: int = 4) -> None:
    """Give the launch the split-K tail scratch (csrc/gemm_sm100.cuh ``sk_*``): the tiles of the last partial wave are cut
    into K ranges that run on the otherwise idle CTA pairs (768 tiles on 74 pairs: 10.5 waves instead of 11).  One scratch
    per (device, stream) -- two GEMMs on different streams never share partial sums.  ``TD_SPLITK=0`` turns it off."""
    import os
    if os.environ.get("TD_SPLITK", "1") == "0":
        return
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    sc = _SK_SCRATCH.get(key)
    if sc is None:
        if torch.cuda.is_current_stream_capturing():
            return                      # never allocate inside a capture; the unsplit schedule is always valid
        sc = (torch.empty(_SK_WS_BYTES, dtype=torch.uint8, device=device), torch.zeros(_SK_FLAGS, dtype=torch.int32, device=device))
        _SK_SCRATCH[key] = sc
    args.sk_ws, args.sk_ws_bytes = sc[0].data_ptr(), _SK_WS_BYTES
    args.sk_flags, args.sk_flag_count, args.sk_max_parts = sc[1].data_ptr(), _SK_FLAGS, max_parts


_C.register("td_gemv", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p])


def gemv(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Skinny GEMM for decode (M <= 8): weight-streaming CUDA-core kernel (csrc/elementwise.cu ``gemv_kernel``)."""
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _C.check(_C.cuda_lib().td_gemv(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), out.stride(0),
                                   int(a.dtype == torch.bfloat16), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemv")
    return out


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
         config: Optional[GemmConfig] = None, out_parity=None, split_k: bool = True) -> torch.Tensor:
    """``out[M,N] = a[M,K] @ b[N,K].T`` with fp32 accumulation in TMEM.  ``b`` is an ``nn.Linear`` weight.

    ``out_parity=(phase_tensor, stride_bytes)``: ``out`` is half 0 of a parity-double-buffered staging area; the
    kernel writes half ``(phase_tensor[0] + 1) & 1`` chosen ON THE DEVICE (CUDA-graph replay safe).
    ``split_k``: allow the split-K tail schedule (see :func:`attach_splitk`)."""
    if not a.is_cuda:
        raise RuntimeError("triton_dist.ops.gemm needs CUDA tensors (sm_100a kernel); the CPU path is emulation-only")
    _check_operand(a, "a")
    _check_operand(b, "b")
    M, K = a.shape
    N, Kb = b.shape
    if K != Kb or a.dtype != b.dtype:
        raise ValueError("shape/dtype mismatch")
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    if (M <= 8 and config is None and out_parity is None and b.is_contiguous() and out.stride(1) == 1 and M * K * 2 <= 200 * 1024
            and a.stride(0) % 8 == 0):
        return gemv(a, b, out)           # decode: tensor cores cannot help at M <= 8, stream the weights instead
    cfg = config or default_config(M, N, K)
    args = _C.GemmArgs()
    args.mode = 0
    fill_common(args, M, a.data_ptr(), a.stride(0), b, out.data_ptr(), M, out.stride(0), M, N, K, cfg,
                a.dtype == torch.bfloat16)
    if out_parity is not None:
        args.c_phase, args.c_nbuf, args.c_buf_stride_bytes = out_parity[0].data_ptr(), 2, int(out_parity[1])
    if split_k:
        attach_splitk(args, a.device)
    lib = _C.cuda_lib()
    _C.check(lib.td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemm_launch")
    return out


_Q8_CODE = {torch.int8: 3}
if hasattr(torch, "float8_e4m3fn"):
    _Q8_CODE[torch.float8_e4m3fn] = 4


def _scale_vec(s, n: int, device) -> Optional[torch.Tensor]:
    """None | python float | 0-d / 1-element tensor (per-tensor) | [n] tensor (per-row / per-channel) -> fp32 [n] or None."""
    if s is None:
        return None
    if not isinstance(s, torch.Tensor):
        s = torch.tensor(float(s), device=device)
    s = s.to(device=device, dtype=torch.float32).reshape(-1)
    if s.numel() == 1:
        s = s.expand(n)
    assert s.numel() == n, f"scale has {s.numel()} entries, expected 1 or {n}"
    return s.contiguous()


def gemm_scaled(a: torch.Tensor, b: torch.Tensor, scale_a=None, scale_b=None, out: Optional[torch.Tensor] = None,
                config: Optional[GemmConfig] = None, out_parity=None, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """8-bit GEMM with dequantisation in the epilogue: ``out[M, N] = (a[M, K] @ b[N, K].T) * scale_a[:, None] * scale_b[None, :]``.
    ``a`` / ``b``: int8 (tcgen05 ``kind::i8``, exact int32 accumulation) or float8_e4m3fn (``kind::f8f6f4``, fp32 accumulation);
    ``scale_a``: per-tensor or per-row [M]; ``scale_b``: per-tensor or per-output-channel [N].  The reference's int8 GEMM+AllReduce
    (kernels/nvidia/gemm_allreduce.py:383-447) and the per-tensor fp8 / int8 gemm_rs dtypes (test_gemm_rs.py:130-145) map here."""
    if not a.is_cuda:
        acc = a.to(torch.float32) @ b.to(torch.float32).t()
        sa, sb = _scale_vec(scale_a, a.shape[0], a.device), _scale_vec(scale_b, b.shape[0], a.device)
        if sa is not None:
            acc = acc * sa[:, None]
        if sb is not None:
            acc = acc * sb[None, :]
        res = acc.to(out_dtype)
        if out is not None:
            out.copy_(res)
            return out
        return res
    assert a.dtype == b.dtype and a.dtype in _Q8_CODE, f"gemm_scaled takes int8 or float8_e4m3fn operands, got {a.dtype}"
    M, K = a.shape
    N, Kb = b.shape
    assert K == Kb and K % 128 == 0 and a.stride(1) == 1 and b.stride(1) == 1 and a.stride(0) % 16 == 0 and b.stride(0) % 16 == 0
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.dtype == torch.bfloat16
    sa, sb = _scale_vec(scale_a, M, a.device), _scale_vec(scale_b, N, a.device)
    cfg = config or default_config(M, N, K)
    args = _C.GemmArgs()
    args.mode = 0
    fill_common(args, M, a.data_ptr(), a.stride(0), b, out.data_ptr(), M, out.stride(0), M, N, K, cfg, True)
    args.is_bf16 = _Q8_CODE[a.dtype]
    args.scale_a = sa.data_ptr() if sa is not None else None
    args.scale_b = sb.data_ptr() if sb is not None else None
    if out_parity is not None:
        args.c_phase, args.c_nbuf, args.c_buf_stride_bytes = out_parity[0].data_ptr(), 2, int(out_parity[1])
    _C.check(_C.cuda_lib().td_gemm_launch(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "td_gemm_launch(8-bit)")
    return out


# ---- the reference's local-GEMM entry points (kernels/nvidia/gemm.py:382-907) ------------------------------------------------------
def get_config_space(persistent: bool = True, device_id: int = 0):
    """Every tile configuration the sm_100a kernel accepts -- the search space of the tuning tools (tools/tune/tune_gemm.py).
    ``persistent=False`` keeps only the configurations that use all SMs with the default grid (the kernel itself is always a
    persistent, warp-specialised tcgen05 kernel; there is no one-tile-per-CTA variant to fall back to)."""
    space = []
    for cta_group in (2, 1):
        for bn in (256, 192, 128, 64, 32):
            for group_m in ((8,) if not persistent else (1, 4, 8, 16)):
                for tma_store in (True, False):
                    space.append(GemmConfig(bn=bn, cta_group=cta_group, group_m=group_m, use_tma_store=tma_store))
    return space


def _as_weight(b: torch.Tensor) -> torch.Tensor:
    """The reference's ``matmul(a, b)`` takes ``b`` as [K, N]; the kernel wants the ``nn.Linear`` layout [N, K] (K-major)."""
    bt = b.t()
    return bt if bt.is_contiguous() else bt.contiguous()


def _cfg(a, b, config, tma_store: bool):
    base = config if isinstance(config, GemmConfig) else default_config(a.shape[0], b.shape[1], a.shape[1])
    return GemmConfig(base.bn, base.cta_group, base.group_m, tma_store, base.num_sms, base.n_comm_ctas)


def matmul(a: torch.Tensor, b: torch.Tensor, config: Optional[GemmConfig] = None) -> torch.Tensor:
    """``a[M, K] @ b[K, N]`` with the direct-store epilogue (TMEM -> registers -> 16-byte global stores)."""
    return gemm(a, _as_weight(b), config=_cfg(a, b, config, False))


def matmul_tma(a: torch.Tensor, b: torch.Tensor, config: Optional[GemmConfig] = None, warp_specialize: bool = True) -> torch.Tensor:
    """``a[M, K] @ b[K, N]`` with the TMA-store epilogue (TMEM -> swizzled shared memory -> ``cp.async.bulk.tensor`` store).
    The kernel is always warp-specialised (TMA / MMA / epilogue warps); ``warp_specialize`` is accepted for call compatibility."""
    return gemm(a, _as_weight(b), config=_cfg(a, b, config, True))


matmul_persistent = matmul                                   # the kernel is persistent in both epilogue flavours
matmul_tma_persistent = matmul_descriptor_persistent = matmul_tma

