"""Kernels / variants written after the round's GPU budget was spent.  They compile for sm_100a and are opt-in in the library;
these tests run only with TD_EXPERIMENTAL=1 so that a default `pytest -m gpu` exercises hardware-validated paths only."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("TD_EXPERIMENTAL", "0") != "1", reason="set TD_EXPERIMENTAL=1")]


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("shape", [(512, 768, 512), (4096, 12288, 1024), (1000, 392, 264)])
def test_gemm_bn192(cg, shape):
    """192-wide tiles (gemm_rs wave quantisation: 4096x12288 -> 13.8 waves of 256x192 instead of 10.4 of 256x256)."""
    from triton_dist.ops.gemm import GemmConfig, gemm
    M, N, K = shape
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
    for tma in (True, False):
        c = gemm(a, b, config=GemmConfig(bn=192, cta_group=cg, group_m=8, use_tma_store=tma))
        torch.testing.assert_close(c.float(), a.float() @ b.float().t(), atol=0.5, rtol=2e-2)


@pytest.mark.parametrize("Sq,Sk,causal", [(256, 256, True), (200, 333, True), (128, 1024, False), (1000, 1000, True), (4096, 4096, True)])
def test_flash_v3(Sq, Sk, causal):
    """Two query tiles per CTA (flash_fwd_kernel_v3) vs the fp32 reference, plus its speed against v2."""
    from triton_dist.ops.flash_attn import flash_attn_fwd, flash_attn_reference
    torch.manual_seed(Sq + Sk)
    q = torch.randn(2, Sq, 8, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(2, Sk, 2, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(2, Sk, 2, 128, device="cuda", dtype=torch.bfloat16)
    out, lse = flash_attn_fwd(q, k, v, causal=causal, return_lse=True, v3=True)
    ref, ref_lse = flash_attn_reference(q, k, v, causal)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(lse, ref_lse, atol=2e-2, rtol=1e-2)


def test_flash_v3_perf():
    from triton_dist.ops.flash_attn import flash_attn_fwd
    q = torch.randn(1, 8192, 32, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(1, 8192, 8, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(1, 8192, 8, 128, device="cuda", dtype=torch.bfloat16)
    for name, kw in (("v2", dict()), ("v3", dict(v3=True))):
        for _ in range(3):
            flash_attn_fwd(q, k, v, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            flash_attn_fwd(q, k, v, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"\nflash {name}: {ms:.3f} ms  {4 * 32 * 8192 * 8192 * 128 / 2 / ms / 1e9:.0f} TFLOP/s")


def test_allreduce_ll_two_gpus():
    """Flag-in-data all-reduce (csrc/allreduce_ll.cu) on 2 GPUs."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _launch import run_dist
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    print(run_dist(["allreduce_ll"], nproc=2, timeout=200)[-600:])
