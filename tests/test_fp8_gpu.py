"""MXFP8 block-scaled GEMM (tcgen05 kind::mxf8f6f4.block_scale) vs the fp32 reference of the dequantised operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_quantizer_matches_host_reference():
    from triton_dist.ops.fp8 import dequantize_mxfp8, quantize_mxfp8
    x = (torch.randn(300, 512, device="cuda") * 4).to(torch.bfloat16)
    t = quantize_mxfp8(x)
    th = quantize_mxfp8(x.cpu().float())
    assert torch.equal(t.sf.cpu(), th.sf)
    d = dequantize_mxfp8(t)
    assert ((d - x.float()).abs().max() / x.float().abs().max()).item() < 0.07


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("shape", [(256, 128, 128), (512, 384, 1024), (300, 264, 640), (4096, 4096, 4096)])
def test_gemm_mxfp8(shape, cta_group, bn):
    from triton_dist.ops.fp8 import dequantize_mxfp8, gemm_mxfp8, quantize_mxfp8
    from triton_dist.ops.gemm import GemmConfig
    torch.manual_seed(0)
    M, N, K = shape
    a = quantize_mxfp8((torch.randn(M, K, device="cuda") * 2).to(torch.bfloat16))
    b = quantize_mxfp8((torch.randn(N, K, device="cuda") * 2).to(torch.bfloat16))
    c = gemm_mxfp8(a, b, config=GemmConfig(bn=bn, cta_group=cta_group, group_m=4, use_tma_store=True))
    torch.cuda.synchronize()
    ref = dequantize_mxfp8(a) @ dequantize_mxfp8(b).t()
    err = (c.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 0.05, (err, ref.abs().max().item())
