"""Test helpers (reference: /root/reference/python/triton_dist/test/utils.py:31-134)."""
import torch

# (name, N, K) of the projection GEMMs swept by the reference benchmarks (test/utils.py:31-39)
LAYER_CONFIGS = {
    "LLaMA-7B": {"N": 11008, "K": 4096, "BM": 128, "BN": 128, "BK": 64},
    "LLaMA-3.1-8B": {"N": 14336, "K": 4096, "BM": 128, "BN": 128, "BK": 64},
    "LLaMA-3.1-70B": {"N": 28672, "K": 8192, "BM": 128, "BN": 256, "BK": 64},
    "LLaMA-3.1-405B": {"N": 53248, "K": 16384, "BM": 128, "BN": 256, "BK": 64},
    "Mistral-7B": {"N": 14336, "K": 4096, "BM": 128, "BN": 128, "BK": 64},
    "Qwen2-72B": {"N": 29568, "K": 8192, "BM": 128, "BN": 256, "BK": 64},
    "GPT-3-175B": {"N": 49152, "K": 12288, "BM": 128, "BN": 256, "BK": 64},
}

THRESHOLD_MAP = {torch.float16: 1e-2, torch.bfloat16: 1e-2, torch.float32: 1e-4, torch.float8_e4m3fn: 2e-2, torch.float8_e5m2: 2e-2, torch.int8: 0}


def assert_allclose(x: torch.Tensor, y: torch.Tensor, atol: float = 1e-3, rtol: float = 1e-3, verbose: bool = True):
    x, y = x.float().cpu(), y.float().cpu()
    if not torch.allclose(x, y, atol=atol, rtol=rtol):
        diff = (x - y).abs()
        bad = diff > (atol + rtol * y.abs())
        msg = f"mismatch: {int(bad.sum())}/{bad.numel()} elements, max abs diff {diff.max().item():.5g}"
        if verbose:
            idx = bad.nonzero()[:8]
            msg += "\n" + "\n".join(f"  at {tuple(i.tolist())}: {x[tuple(i)].item():.6g} vs {y[tuple(i)].item():.6g}" for i in idx)
        raise AssertionError(msg)


def assert_bitwise_equal(x: torch.Tensor, y: torch.Tensor):
    xb, yb = x.contiguous().view(torch.uint8).cpu(), y.contiguous().view(torch.uint8).cpu()
    if xb.shape != yb.shape or not torch.equal(xb, yb):
        n = int((xb != yb).sum()) if xb.shape == yb.shape else -1
        raise AssertionError(f"tensors differ bitwise ({n} bytes)")


def bitwise_equal(x: torch.Tensor, y: torch.Tensor) -> bool:
    """True when the two tensors have identical bytes (NaN == NaN, -0.0 != +0.0)."""
    if x.shape != y.shape or x.dtype != y.dtype:
        return False
    return bool(torch.equal(x.contiguous().view(torch.uint8), y.contiguous().view(torch.uint8)))
