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
