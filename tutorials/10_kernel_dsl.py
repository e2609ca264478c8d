"""Tutorial 10 -- kernels in Python: the ``triton_dist.lk`` DSL (reference: python/little_kernel and its gemm_sm100 ladder).

A DSL kernel is a typed Python function.  The same source (1) lowers to CUDA C++ against the device headers of this framework and is
compiled by nvcc for sm_100a, (2) runs on CPU tensors in the interpreter (threads emulated, ``symm_at`` / ``notify`` / ``wait`` on the
shared-memory heap) -- which is how this tutorial runs without GPUs:
    bash scripts/launch.sh --nproc_per_node=2 tutorials/10_kernel_dsl.py            (GPUs)
    TD_FORCE_HOST_BACKEND=1 bash scripts/launch.sh --nproc_per_node=2 tutorials/10_kernel_dsl.py
"""
import torch
import triton_dist.utils as U
from triton_dist import lk
from triton_dist.lk import ll

U.initialize_distributed(seed=0)
W, me, dev = U.world_size(), U.rank(), U.current_device()
gpu = dev.type == "cuda"

# ---- 1. a SIMT kernel: Python globals are compile-time constants, helpers become inlined __device__ functions -----------------
BLOCK = 128


def warp_sum(v):                                   # specialised per argument type, like a C++ template
    for off in ll.static_range(4, -1, -1):         # unrolled by the code generator
        v = v + ll.shfl_xor(v, 1 << off)
    return v


@lk.kernel(block=BLOCK)
def row_norm(x: ll.ptr[ll.f32], y: ll.ptr[ll.f32], cols: ll.i32, eps: ll.f32):
    """y[r, :] = x[r, :] * rsqrt(mean(x[r, :]^2) + eps) -- one block per row."""
    part = ll.shared([BLOCK // 32], ll.f32)
    tid = ll.threadIdx.x
    row = x + ll.i64(ll.blockIdx.x) * cols
    acc: ll.f32 = 0.0
    for c in range(tid, cols, BLOCK):
        acc += row[c] * row[c]
    acc = warp_sum(acc)
    if tid % 32 == 0:
        part[tid // 32] = acc
    ll.syncthreads()
    total: ll.f32 = 0.0
    for w in ll.static_range(BLOCK // 32):
        total += part[w]
    scale = ll.rsqrt(total / cols + eps)
    out = y + ll.i64(ll.blockIdx.x) * cols
    for c in range(tid, cols, BLOCK):
        out[c] = row[c] * scale


x = torch.randn(6, 500, device=dev)
y = torch.empty_like(x)
if gpu:
    row_norm[6](x, y, 500, 1e-6)                   # nvcc (cached by content hash) + cudaLaunchKernelEx on the current stream
else:
    row_norm.interpret(6, x, y, 500, 1e-6)         # the same Python source, threads emulated
ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)
U.dist_print(f"row_norm: max err {(y - ref).abs().max().item():.2e}; generated C++ is {len(row_norm.cuda_source().splitlines())} lines",
             allowed_ranks=[0])
assert torch.allclose(y, ref, atol=1e-5)

# ---- 2. a distributed kernel: the symmetric-heap primitives are DSL intrinsics ----------------------------------------------------


@lk.kernel(block=BLOCK)
def exchange(ctx: ll.SymmCtx, src: ll.ptr[ll.f32], inbox: ll.ptr[ll.f32], flag: ll.ptr[ll.u32], n: ll.i32, phase: ll.u32):
    """Write my vector into the inbox of the next rank, raise its flag (release), wait for mine (acquire)."""
    nxt = (ll.rank(ctx) + 1) % ll.num_ranks(ctx)
    remote = ll.symm_at(ctx, inbox, nxt)           # same offset in the peer's heap segment (an NVLink store on a B200)
    for i in range(ll.threadIdx.x, n, BLOCK):
        remote[i] = src[i]
    ll.syncthreads()
    if ll.threadIdx.x == 0:
        ll.notify(ctx, flag, nxt, phase)
    if ll.threadIdx.x < 32:
        ll.wait(flag, 1, phase)
    ll.syncthreads()


n = 256
inbox = U.nvshmem_create_tensor((n,), torch.float32)
flag = U.nvshmem_create_tensor((1,), torch.int32)
flag.zero_()
U.barrier_all_on_stream()
ctx = lk.symm_ctx()
for phase in (1, 2, 3):                            # phase numbers instead of flag resets: replayable, no extra barrier
    src = torch.full((n,), float(100 * me + phase), device=dev)
    if gpu:
        exchange[1](ctx, src, inbox, flag, n, phase)
    else:
        exchange.interpret(1, ctx, src, inbox, flag, n, phase)
    prev = (me - 1 + W) % W
    assert torch.all(inbox == float(100 * prev + phase)), (me, phase, inbox[:4])
    U.barrier_all_on_stream()
U.dist_print("exchange: 3 phases OK", allowed_ranks=[0])

# ---- 3. tensor cores: the tcgen05 GEMM ladder is written in the same DSL -----------------------------------------------------------
# triton_dist/lk/kernels/gemm_sm100.py: TMA -> 4-stage smem ring -> single-thread tcgen05.mma into TMEM -> tcgen05.ld epilogue, as a
# 1-CTA kernel and as a cta_group::2 cluster kernel.  Without a GPU we can still generate the code and (with nvcc) cross-compile it.
from triton_dist.lk.kernels.gemm_sm100 import get_gemm, run_gemm  # noqa: E402

g2 = get_gemm(BN=256, STAGES=4, cta_group=2)
src_txt = g2.cuda_source()
assert "td::ptx::mma_f16<2>" in src_txt and "td::ptx::tma_load_2d_2sm" in src_txt
U.dist_print(f"gemm ladder (cta_group::2): {g2.dyn_smem_bytes} B dynamic smem, tile {g2.tile[:3]}", allowed_ranks=[0])
if gpu:
    a = (torch.randn(1024, 512, device=dev) * 0.5).bfloat16()
    b = (torch.randn(768, 512, device=dev) * 0.5).bfloat16()
    c = run_gemm(a, b, cta_group=2)
    err = (c.float() - a.float() @ b.float().t()).abs().max().item()
    U.dist_print(f"lk tcgen05 gemm: max err {err:.3f}", allowed_ranks=[0])
    assert err < 0.5
# ---- 4. compute + communication in one Python function --------------------------------------------------------------------------
# lk/kernels/ag_gemm.py (comm CTAs push shards + release-add flags, tcgen05 tiles acquire them) and lk/kernels/gemm_rs.py (tile
# epilogues reduce into the owner over NVLink, collector CTAs acquire the arrival counter) are the two halves of a tensor-parallel MLP
# as single DSL kernels.  On GPUs this runs them; on the emulation backend the interpreter's pipeline model does (slow: tiny shapes).
if gpu or W == 2:
    from triton_dist.lk.kernels.ag_gemm import LkAgGemmContext, run_ag_gemm  # noqa: E402
    Ms, K2, N2 = (256, 512, 768) if gpu else (128, 64, 128)
    agc = LkAgGemmContext(Ms, K2, BN=128, STAGES=2, N_COMM=2)
    a_sh = (torch.randn(Ms, K2, device=dev) * 0.5).bfloat16()
    b_w = (torch.randn(N2, K2, generator=torch.Generator().manual_seed(5)) * 0.5).bfloat16().to(dev)
    y = run_ag_gemm(agc, a_sh, b_w)
    full_a = torch.empty(W * Ms, K2, dtype=torch.bfloat16, device=dev)
    torch.distributed.all_gather_into_tensor(full_a, a_sh, group=U.get_triton_dist_world())
    err2 = (y.float() - full_a.float() @ b_w.float().t()).abs().max().item()
    U.dist_print(f"lk ag_gemm (one DSL kernel, {agc.kernel.n_comm} comm CTAs): max err {err2:.3f}", allowed_ranks=[0])
    assert err2 < 0.3
    agc.finalize()
U.barrier_all_host()
U.nvshmem_free_tensor_sync(flag)
U.nvshmem_free_tensor_sync(inbox)
U.dist_print("kernel DSL tutorial OK", allowed_ranks=[0])
U.finalize_distributed()
