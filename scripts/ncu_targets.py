"""One call of one kernel family (after warm-up) for `ncu -k regex:... -s N -c 1`; see scripts/gpu_ncu.sh."""
import sys
import torch
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.ops import GemmConfig, gemm

t = sys.argv[1]
U.initialize_distributed(seed=0)
dev, bf = "cuda", torch.bfloat16
if t in ("gemm_k49152", "gemm_4096", "gemm_rs_shape"):
    M, N, K = {"gemm_k49152": (4096, 12288, 49152), "gemm_4096": (4096, 4096, 4096), "gemm_rs_shape": (4096, 12288, 6144)}[t]
    a = torch.randn(M, K, device=dev, dtype=bf) * 0.05
    b = torch.randn(N, K, device=dev, dtype=bf) * 0.05
    for _ in range(4):
        gemm(a, b)
elif t == "mxfp8":
    from triton_dist.ops.fp8 import gemm_mxfp8, quantize_mxfp8
    a = quantize_mxfp8(torch.randn(4096, 6144, device=dev, dtype=bf) * 0.05)
    b = quantize_mxfp8(torch.randn(12288, 6144, device=dev, dtype=bf) * 0.05)
    for _ in range(4):
        gemm_mxfp8(a, b, config=GemmConfig(256, 2, 8, True))
elif t == "flash":
    from triton_dist.ops.flash_attn import flash_attn_fwd
    q = torch.randn(1, 8192, 32, 128, device=dev, dtype=bf)
    k = torch.randn(1, 8192, 8, 128, device=dev, dtype=bf)
    v = torch.randn(1, 8192, 8, 128, device=dev, dtype=bf)
    for _ in range(4):
        flash_attn_fwd(q, k, v, causal=True)
elif t == "grouped":
    from triton_dist.ops import moe as M_
    x = (torch.randn(16384, 1792, device=dev) * 0.3).to(bf)
    w = (torch.randn(8, 4096, 1792, device=dev) * 0.05).to(bf)
    ids = torch.rand(8192, 8, device=dev).topk(2, dim=1).indices.to(torch.int32)
    for _ in range(4):
        M_.moe_grouped_gemm_presorted(x, w, ids, 8, 1)
elif t == "decode":
    from triton_dist.ops.flash_decode import gqa_fwd_batch_decode
    kc = torch.randn(8, 8192, 8, 128, device=dev, dtype=bf)
    vc = torch.randn(8, 8192, 8, 128, device=dev, dtype=bf)
    q = torch.randn(8, 32, 128, device=dev, dtype=bf)
    lens = torch.full((8,), 8192, device=dev, dtype=torch.int32)
    for _ in range(4):
        gqa_fwd_batch_decode(q, kc, vc, lens)
elif t == "gemv":
    a = torch.randn(4, 4096, device=dev, dtype=bf)
    b = torch.randn(12288, 4096, device=dev, dtype=bf)
    for _ in range(4):
        gemm(a, b)
elif t == "mega":
    from triton_dist.mega_kernel import MegaDenseModel
    from triton_dist.models import AutoLLM, KV_Cache, ModelConfig
    cfg = ModelConfig(model_name="Qwen/Qwen3-8B", max_length=1024, dtype=bf, rank=0, world_size=1, num_layers_override=4)
    m = AutoLLM.from_pretrained(cfg)
    kv = KV_Cache(m.num_layers, 1, 1024, m.num_key_value_heads, m.head_dim, bf, 1, dev)
    kv.rand_fill_kv_cache(512)
    mega = MegaDenseModel(m, 1, kv)
    tok = torch.randint(0, 1000, (1, 1), device=dev)
    for _ in range(4):
        mega.mega_forward(tok)
torch.cuda.synchronize()
