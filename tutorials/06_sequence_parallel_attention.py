"""Tutorial 06 -- long-context attention: the tcgen05 flash-attention kernel, context parallelism over an all-gathered KV with
zig-zag sharding, Ulysses head<->sequence all-to-all fused into the QKV GEMM, and KV-sharded flash-decode (reference:
sp_ag_attention_intra_node.py, ulysses_sp_dispatch.py, sp_flash_decode_layer.py).
    bash scripts/launch.sh --nproc_per_node=2 tutorials/06_sequence_parallel_attention.py"""
import math
import torch
import triton_dist.utils as U
from triton_dist.ops.flash_attn import flash_attn_fwd, flash_attn_reference
from triton_dist.ops.gemm_a2a import create_gemm_a2a_context, gemm_all_to_all
from triton_dist.parallel.sp import create_sp_ag_attention_context_intra_node, fused_sp_ag_attn_intra_node, zigzag_positions

U.initialize_distributed(seed=0)
W, me, dev = U.world_size(), U.rank(), U.current_device()
gpu = dev.type == "cuda"
dt = torch.bfloat16 if gpu else torch.float32
Hq, Hkv, D = 4, 2, 128

# 1. one device: causal GQA attention (tcgen05 kernel on a B200, fp32 reference on the emulation backend)
S = 1024 if gpu else 64
q, k, v = (torch.randn(1, S, h, D, device=dev).to(dt) for h in (Hq, Hkv, Hkv))
o = flash_attn_fwd(q, k, v, causal=True)
ref, _ = flash_attn_reference(q, k, v, True)
U.dist_print(f"flash_attn_fwd max err {(o.float() - ref).abs().max().item():.3e}", allowed_ranks=[0])

# 2. context parallel prefill: every rank holds a zig-zag shard of the sequence, K/V are all-gathered over NVLink
S = (512 if gpu else 8) * W
g = torch.Generator().manual_seed(1)
qf, kf, vf = (torch.randn(S, h, D, generator=g).to(dt).to(dev) for h in (Hq, Hkv, Hkv))
pos = zigzag_positions(S, W, me, dev) if W > 1 else torch.arange(S, device=dev)
ctx = create_sp_ag_attention_context_intra_node(S // W, Hkv, D, dt)
o = fused_sp_ag_attn_intra_node(ctx, qf[pos].contiguous(), kf[pos].contiguous(), vf[pos].contiguous(), is_causal=True)
full, _ = flash_attn_reference(qf[None], kf[None], vf[None], True)
U.dist_print(f"context-parallel attention max err {(o.float() - full[0][pos]).abs().max().item():.3e}", allowed_ranks=[0])
ctx.finalize()

# 3. Ulysses: QKV projection whose epilogue scatters each rank's heads directly into that rank's buffer
rows, K, c = (256 if gpu else 8), (512 if gpu else 32), (256 if gpu else 32)
x = (torch.randn(rows, K, device=dev) * 0.3).to(dt)
w = (torch.randn(W * c, K, generator=torch.Generator().manual_seed(2)) * 0.3).to(dt).to(dev)      # rows grouped by destination rank
a2a = create_gemm_a2a_context(rows, c, dt)
mine = gemm_all_to_all(a2a, x, w)                      # [W * rows, c]: block s = rank s's tokens x my heads
U.dist_print(f"gemm + all-to-all: got {tuple(mine.shape)}, own block err "
             f"{(mine[me * rows:(me + 1) * rows].float() - x.float() @ w[me * c:(me + 1) * c].float().t()).abs().max().item():.3e}")
a2a.finalize()
U.finalize_distributed()
