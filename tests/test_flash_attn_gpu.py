"""tcgen05 flash-attention forward vs a plain PyTorch fp32 reference (real B200)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(B, Sq, Sk, Hq, Hkv, causal, dtype, cache_len=None, tile_pos=False):
    for bn in (64, 128, 129, 130):     # 129 = 128-key tiles with Q / P operands in TMEM, 130 = v2 kernel
        _check_bn(B, Sq, Sk, Hq, Hkv, causal, dtype, cache_len, tile_pos, bn)


def _check_bn(B, Sq, Sk, Hq, Hkv, causal, dtype, cache_len, tile_pos, bn):
    from triton_dist.ops.flash_attn import flash_attn_fwd, flash_attn_reference
    torch.manual_seed(Sq * 7 + Sk)
    L = cache_len or Sk
    q = torch.randn(B, Sq, Hq, 128, device="cuda", dtype=dtype)
    k = torch.randn(B, L, Hkv, 128, device="cuda", dtype=dtype)
    v = torch.randn(B, L, Hkv, 128, device="cuda", dtype=dtype)
    q_tile_pos = q_pos = None
    if tile_pos:       # every 128-query tile sits at its own place in the KV sequence (zig-zag style)
        nt = (Sq + 127) // 128
        starts = torch.randint(0, max(1, (Sk - 128) // 64), (B, nt), device="cuda", dtype=torch.int32) * 64
        q_tile_pos = starts.contiguous()
        q_pos = (starts.long()[:, :, None] + torch.arange(128, device="cuda")[None, None]).reshape(B, -1)[:, :Sq]
    out, lse = flash_attn_fwd(q, k, v, causal=causal, q_tile_pos=q_tile_pos, sk=Sk, return_lse=True, block_n=min(bn, 128), tmem_operands=(bn == 129), v2=(bn == 130))
    ref, ref_lse = flash_attn_reference(q, k, v, causal, None, q_pos, Sk)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(lse, ref_lse, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [True, False])
def test_flash_square(dtype, causal):
    _check(1, 256, 256, 4, 2, causal, dtype)


@pytest.mark.parametrize("Sq,Sk", [(128, 128), (200, 333), (128, 1024), (1000, 1000), (77, 4096)])
def test_flash_ragged(Sq, Sk):
    _check(2, Sq, Sk, 8, 2, True, torch.bfloat16)
    _check(2, Sq, Sk, 8, 8, False, torch.bfloat16)


def test_flash_kv_cache_bound():
    _check(2, 300, 700, 4, 1, True, torch.bfloat16, cache_len=1024)


def test_flash_tile_positions():
    _check(2, 512, 2048, 4, 2, True, torch.bfloat16, tile_pos=True)


def test_flash_long():
    _check(1, 4096, 4096, 2, 1, True, torch.bfloat16)


def test_flash_perf():
    """Device time of causal 8K x 8K, 32 heads: prints TFLOP/s (informational) and asserts it is a tensor-core number."""
    from triton_dist.ops.flash_attn import flash_attn_fwd
    B, S, H = 1, 8192, 32
    q = torch.randn(B, S, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, S, 8, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, S, 8, 128, device="cuda", dtype=torch.bfloat16)
    tf = 0.0
    for bn in (64, 128, 129, 130):
        for _ in range(3):
            flash_attn_fwd(q, k, v, block_n=min(bn, 128), tmem_operands=(bn == 129), v2=(bn == 130))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            flash_attn_fwd(q, k, v, block_n=min(bn, 128), tmem_operands=(bn == 129), v2=(bn == 130))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t = 4 * B * H * S * S * 128 / 2 / ms / 1e9
        tf = max(tf, t)
        print(f"\nflash_attn causal 8Kx8K 32h block_n={bn}: {ms:.3f} ms  {t:.0f} TFLOP/s")
    try:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        qq, kk, vv = q.transpose(1, 2), k.transpose(1, 2).repeat_interleave(4, 1), v.transpose(1, 2).repeat_interleave(4, 1)
        for _ in range(3):
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True)
        t0.record()
        for _ in range(10):
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True)
        t1.record()
        torch.cuda.synchronize()
        print(f"torch SDPA same problem: {t0.elapsed_time(t1) / 10:.3f} ms")
    except Exception as e:      # noqa: BLE001
        print("sdpa baseline unavailable:", e)
    assert tf > 100
