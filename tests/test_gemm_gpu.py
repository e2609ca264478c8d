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


@pytest.mark.parametrize("cfg", [(256, 2), (256, 1), (128, 2), (128, 1), (64, 1), (192, 2)])
@pytest.mark.parametrize("shape", [(4096, 4096, 4096), (2048, 2560, 2048), (1024, 1024, 8192), (640, 1280, 3072)])
@pytest.mark.parametrize("tma_store", [True, False])
def test_gemm_splitk_tail(cfg, shape, tma_store):
    """The split-K tail schedule (last partial wave cut into K ranges, fp32 partials through the scratch, flags re-armed by the
    kernel) must be bit-compatible in structure with the unsplit schedule: compare both against fp32, and run each shape
    twice back to back so a flag that was not re-armed would show."""
    from triton_dist.ops import GemmConfig, gemm
    torch.manual_seed(2)
    M, N, K = shape
    bn, cg = cfg
    if N % bn:
        pytest.skip("N not a multiple of the tile width")
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16) * 0.5
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.5
    ref = _ref(a, b)
    c0 = gemm(a, b, config=GemmConfig(bn=bn, cta_group=cg, group_m=8, use_tma_store=tma_store), split_k=False)
    for _ in range(2):
        c1 = gemm(a, b, config=GemmConfig(bn=bn, cta_group=cg, group_m=8, use_tma_store=tma_store), split_k=True)
        torch.cuda.synchronize()
        tol = dict(atol=0.05 * (K ** 0.5) * 0.25 + 0.05, rtol=2e-2)
        torch.testing.assert_close(c1.float(), ref, **tol)
        torch.testing.assert_close(c0.float(), ref, **tol)
        # same fp32 sums up to the association order of the K split: differences are bf16 rounding flips only
        assert (c1.float() - c0.float()).abs().max().item() <= 0.02 * ref.abs().max().item() + 1e-2


@pytest.mark.parametrize("kind", ["int8", "fp8"])
@pytest.mark.parametrize("shape", [(512, 768, 1024), (300, 264, 512), (4096, 4096, 4096)])
def test_gemm_scaled_8bit(kind, shape):
    """int8 x int8 (tcgen05 kind::i8, exact int32 accumulation) and e4m3 x e4m3 (kind::f8f6f4) with per-row / per-channel
    dequantisation scales applied in the epilogue, vs the fp32 product of the dequantised operands."""
    from triton_dist.ops.gemm import gemm_scaled
    torch.manual_seed(6)
    M, N, K = shape
    if kind == "int8":
        a = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8)
        b = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8)
    else:
        a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.float8_e4m3fn)
        b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.float8_e4m3fn)
    sa = torch.rand(M, device="cuda") * 0.02 + 0.001
    sb = torch.rand(N, device="cuda") * 0.02 + 0.001
    ref = (a.float() @ b.float().t()) * sa[:, None] * sb[None, :]
    out = gemm_scaled(a, b, sa, sb)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float(), ref, atol=2e-2 * ref.abs().max().item() + 1e-3, rtol=2e-2)
    out2 = gemm_scaled(a, b, 0.01, None)                      # per-tensor scale, no weight scale
    torch.testing.assert_close(out2.float(), (a.float() @ b.float().t()) * 0.01, atol=2e-2 * (ref.abs().max().item() / 0.01 * 0.01 * 100) + 1e-2, rtol=2e-2)
