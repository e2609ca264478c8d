"""Numerics of the tcgen05 GEMM vs a plain PyTorch fp32 reference (runs on a real B200)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b):
    return a.float() @ b.float().t()


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("bn", [32, 64, 128, 256])
@pytest.mark.parametrize("tma_store", [False, True])
def test_gemm_configs(bn, cta_group, tma_store):
    from triton_dist.ops import GemmConfig, gemm
    torch.manual_seed(0)
    M, N, K = 512, 768, 512
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    c = gemm(a, b, config=GemmConfig(bn=bn, cta_group=cta_group, group_m=2, use_tma_store=tma_store))
    torch.cuda.synchronize()
    ref = _ref(a, b)
    torch.testing.assert_close(c.float(), ref, atol=0.5, rtol=2e-2)


@pytest.mark.parametrize("shape", [(128, 128, 64), (1, 4096, 4096), (77, 1000, 520), (4096, 4096, 4096), (300, 264, 1032)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_shapes(shape, dtype):
    from triton_dist.ops import gemm
    torch.manual_seed(1)
    M, N, K = shape
    a = torch.randn(M, K, device="cuda", dtype=dtype) * 0.5
    b = torch.randn(N, K, device="cuda", dtype=dtype) * 0.5
    c = gemm(a, b)
    torch.cuda.synchronize()
    ref = _ref(a, b)
    torch.testing.assert_close(c.float(), ref, atol=0.05 * (K ** 0.5) * 0.25 + 0.05, rtol=2e-2)
