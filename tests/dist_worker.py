"""Multi-process test worker (launched by torchrun from tests/test_dist_*.py, on CPU/gloo or on GPUs).

Pattern taken from the reference's test strategy: fresh random inputs every iteration,
poisoned workspaces, golden = torch.distributed collective + matmul, per-rank asserts, straggler injection.
Usage: torchrun ... tests/dist_worker.py <case> [args]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import triton_dist.utils as U  # noqa: E402


def _assert_close(a, b, atol, rtol, what):
    a, b = a.float().cpu(), b.float().cpu()
    if not torch.allclose(a, b, atol=atol, rtol=rtol):
        err = (a - b).abs()
        raise AssertionError(f"[rank {U.rank()}] {what}: max abs err {err.max().item():.4g} at {err.argmax().item()} "
                             f"(ref max {b.abs().max().item():.4g})")


def case_primitives():
    """notify/wait ring + symm_at + barrier (BASELINE config #1; reference tutorials/01, test_notify.py)."""
    import ctypes
    from triton_dist import language as dl
    heap = U.get_heap()
    W, me = U.world_size(), U.rank()
    data = U.nvshmem_create_tensor((64,), torch.float32)
    sig = U.nvshmem_create_tensor((8,), torch.int32)
    U.barrier_all_on_stream()
    nxt = (me + 1) % W
    for rnd in range(1, 6):
        payload = torch.full((64,), float(me * 100 + rnd), dtype=torch.float32, device=data.device)
        # producer: data into the next rank's buffer, then the flag (release)
        U.symm_at(data, nxt).copy_(payload)
        dl.notify(sig[0:1], nxt, signal=rnd, sig_op="set", comm_scope="intra_node")
        # consumer: wait for my predecessor's flag (acquire), then read
        tok = dl.wait(sig[0:1], 1, "sys", "acquire", wait_value=rnd)
        got = dl.consume_token(data, tok).clone()
        prev = (me - 1 + W) % W
        assert torch.all(got == float(prev * 100 + rnd)), (me, rnd, got[:4])
        U.barrier_all_on_stream()
    # ADD signal from every rank onto rank 0 (test_notify.py:57-69)
    dl.notify(sig[1:2], 0, signal=1, sig_op="add")
    U.barrier_all_on_stream()
    if me == 0:
        if data.is_cuda:
            torch.cuda.synchronize()
        assert int(sig[1].item()) == W, sig
    U.barrier_all_on_stream()
    U.nvshmem_free_tensor_sync(sig)
    U.nvshmem_free_tensor_sync(data)


def case_allgather():
    from triton_dist.ops import comm
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    ctx = comm.create_fast_allgather_context(1 << 20)
    modes = ["pull", "push", "push_2d_ll"] if dev.type == "cuda" else ["push"]
    for it in range(4):
        for mode in modes:
            for n in (16, 1000, 65536):
                x = torch.randn(n, device=dev).to(torch.bfloat16 if n % 8 == 0 else torch.float32)
                out = comm.fast_allgather(x, ctx, mode=mode)
                ref = torch.empty(W * x.numel(), dtype=x.dtype, device=dev)
                dist.all_gather_into_tensor(ref, x, group=U.get_triton_dist_world())
                assert torch.equal(out.cpu().view(-1), ref.cpu()), (mode, n, it)
    # intra-node copy-engine allgather (reference contract: allgather.py:100-124)
    from triton_dist.ops.allgather import cp_engine_producer_all_gather_intra_node, create_allgather_buffers
    bufs, flags = create_allgather_buffers(64 * W, 32, torch.float32)
    for it in range(1, 4):
        local = torch.randn(64, 32, device=dev)
        cp_engine_producer_all_gather_intra_node(me, W, local, bufs, flags, signal_value=it)
        U.barrier_all_on_stream()
        ref = torch.empty(64 * W * 32, device=dev)
        dist.all_gather_into_tensor(ref, local.view(-1), group=U.get_triton_dist_world())
        assert torch.equal(bufs[me].cpu().view(-1), ref.cpu())
        assert torch.all(flags[me][:W].cpu() == it)
        U.barrier_all_on_stream()
    ctx.finalize()


def case_allgather_mc():
    """NVLS variants of the fast all-gather (multimem push + barrier, multimem LL atoms), interleaved with the unicast modes on the
    same context so parity buffers / phase counters are shared (reference: low_latency_allgather.py:623-700).  Without the multicast
    mapping (emulation backend, or a driver without NVLS) the names resolve to their unicast twins."""
    from triton_dist.ops import comm
    W = U.world_size()
    dev = U.current_device()
    ctx = comm.create_fast_allgather_context(1 << 20)
    for it in range(4):
        for mode in ("push_multimem", "push_2d_ll_multimem", "push", "ll_multimem", "push_2d_multimem"):
            for n in (16, 1000, 4100, 65536):
                x = torch.randn(n, device=dev).to(torch.bfloat16 if n % 8 == 0 else torch.float32)
                out = comm.fast_allgather(x, ctx, mode=mode)
                ref = torch.empty(W * x.numel(), dtype=x.dtype, device=dev)
                dist.all_gather_into_tensor(ref, x, group=U.get_triton_dist_world())
                assert torch.equal(out.cpu().view(-1), ref.cpu()), (mode, n, it)
    ctx.finalize()
    from triton_dist.layers.nvidia import AllGatherLayer
    layer = AllGatherLayer(1 << 18)
    for name in ("forward_pull", "forward_push_2d", "forward_push_3d", "forward_push_2d_ll", "forward_push_numa_2d", "forward_push_numa_2d_ll",
                 "forward_push_multimem", "forward_push_2d_ll_multimem", "forward"):
        if dev.type != "cuda" and name == "forward_pull":
            continue
        for n in (64, 30000):
            x = torch.randn(n, device=dev)
            out = getattr(layer, name)(x)
            ref = torch.empty(W * n, device=dev)
            dist.all_gather_into_tensor(ref, x, group=U.get_triton_dist_world())
            assert torch.equal(out.cpu().view(-1), ref.cpu()), (name, n)
    layer.finalize()


def case_allreduce():
    from triton_dist.ops import comm
    dev = U.current_device()
    ctx = comm.create_allreduce_ctx(4 << 20, U.rank(), U.world_size(), U.world_size())
    methods = [comm.AllReduceMethod.OneShot, comm.AllReduceMethod.TwoShot]
    if U.is_nvshmem_multimem_supported():
        methods += [comm.AllReduceMethod.OneShot_Multimem, comm.AllReduceMethod.TwoShot_Multimem]
    if dev.type != "cuda":
        methods = [comm.AllReduceMethod.OneShot]
    U.dist_print(f"allreduce methods: {[m.name for m in methods]} multimem={U.is_nvshmem_multimem_supported()}", allowed_ranks=[0])
    gen = torch.Generator().manual_seed(1234)
    for it in range(6):
        for dtype in (torch.bfloat16, torch.float32, torch.float16):
            n = int(torch.randint(1, 300000, (1,), generator=gen).item()) * 8
            if it == 0:
                n = 8
            if it == 5:
                n = (6 << 20) // 2     # larger than the workspace -> chunked
            for m in methods:
                x = (torch.randn(n, device=dev) * 0.5).to(dtype)
                ref = x.clone()
                dist.all_reduce(ref, group=U.get_triton_dist_world())
                straggler = (it % U.world_size(), 2_000_000) if (dev.type == "cuda" and it in (2, 3)) else None
                out = comm.all_reduce(x, m, ctx, straggler_option=straggler)
                tol = 2e-2 if dtype != torch.float32 else 1e-4
                _assert_close(out, ref, tol * 4, tol, f"allreduce {m.name} {dtype} n={n}")
    ctx.finalize()


def case_allreduce_ll():
    """Flag-in-data low-latency all-reduce (opt-in; GPU only): numerics over many back-to-back calls (buffer halves and phases are
    reused), with a straggler, plus its CUDA-graph latency next to the NVLS one-shot kernel."""
    from triton_dist.ops import comm
    dev = U.current_device()
    if dev.type != "cuda":
        return
    W = U.world_size()
    ctx = comm.create_allreduce_ctx(1 << 20, U.rank(), W, W)
    gen = torch.Generator().manual_seed(77)
    for it in range(40):
        n = int(torch.randint(1, 2048, (1,), generator=gen).item()) * 8
        for dtype in (torch.bfloat16, torch.float16, torch.float32):
            x = (torch.randn(n, device=dev) * 0.5).to(dtype)
            ref = x.clone()
            dist.all_reduce(ref, group=U.get_triton_dist_world())
            if it in (3, 17) and U.rank() == it % W:
                torch.cuda._sleep(2_000_000)
            out = comm.all_reduce(x, comm.AllReduceMethod.OneShot_LL, ctx)
            tol = 2e-2 if dtype != torch.float32 else 1e-4
            _assert_close(out, ref, tol * 4, tol, f"allreduce LL {dtype} n={n} it{it}")
    x = torch.randn(4096, device=dev).to(torch.bfloat16)
    o = torch.empty_like(x)
    for name, m in (("OneShot_LL", comm.AllReduceMethod.OneShot_LL), ("OneShot_Multimem", comm.AllReduceMethod.OneShot_Multimem)):
        for _ in range(3):
            comm.all_reduce(x, m, ctx, output=o)
        torch.cuda.synchronize(); dist.barrier(group=U.get_triton_dist_world())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                comm.all_reduce(x, m, ctx, output=o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay(); torch.cuda.synchronize(); dist.barrier(group=U.get_triton_dist_world())
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        U.dist_print(f"all_reduce 8 KB {name}: {e0.elapsed_time(e1) / 200 * 1e3:.2f} us per call (CUDA graph)", allowed_ranks=[0])
    ctx.finalize()


def case_ag_gemm():
    from triton_dist.ops.ag_gemm import ag_gemm, create_ag_gemm_context
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    shapes = [(512 * W, 512, 1024), (256 * W, 256, 512), (384 * W, 768, 256), (100 * W, 264, 520), (512 * W, 512, 4096),
              (128 * W, 1280, 2112)] if big else [(16 * W, 24, 32), (5 * W, 8, 16)]
    dtype = torch.bfloat16 if big else torch.float32
    transports = (["sm", "sm_k"] + (["multicast"] if U.is_nvshmem_multimem_supported() else [])) if big else ["auto"]
    for (M, N, K) in shapes:
        ctx = create_ag_gemm_context(M, N, K, dtype)
        if big:
            ctx.workspace.view(torch.int16).fill_(0x7FC0)      # poison (bf16 NaN pattern)
        for tr in transports:
            if tr in ("multicast", "sm_k") and (M // W) % 128 != 0:
                continue
            for it in range(5):
                A = (torch.randn(M // W, K, device=dev) * 0.5).to(dtype)
                Wt = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
                straggler = (it % W, 3_000_000) if (big and it in (1, 3)) else None
                ks = (0, 4, 16, 3, 8)[it] if tr in ("multicast", "sm_k") else 0
                C = ag_gemm(A, Wt.t(), ctx, straggler_option=straggler, transport=tr, kslices=ks, comm_groups=(0, 2, 4, 1, 3)[it], tail_pct=(0, 10, 25, 0, 12)[it])
                full = torch.empty(M * K, device=dev, dtype=dtype)
                dist.all_gather_into_tensor(full, A.view(-1), group=U.get_triton_dist_world())
                ref = full.view(M, K).float() @ Wt.float().t()
                _assert_close(C, ref, 0.5 if big else 1e-3, 2e-2 if big else 1e-4, f"ag_gemm[{tr}] {M}x{N}x{K} it{it}")
        # AllToAll + GEMM in the same kernel (all_to_all_single_gemm.py:74-188): block d of x goes to rank d
        from triton_dist.ops.compat import all_to_all_single_gemm
        for it in range(3):
            X = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
            Wt = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
            C = all_to_all_single_gemm(ctx, X, Wt)
            recv = torch.empty_like(X)
            if dev.type == "cuda":
                dist.all_to_all_single(recv, X, group=U.get_triton_dist_world())
            else:
                allx = torch.empty(W * M * K, dtype=dtype)
                dist.all_gather_into_tensor(allx, X.view(-1), group=U.get_triton_dist_world())
                Ms = M // W
                recv = torch.cat([allx.view(W, M, K)[s, me * Ms:(me + 1) * Ms] for s in range(W)])
            _assert_close(C, recv.float() @ Wt.float().t(), 0.5 if big else 1e-3, 2e-2 if big else 1e-4, f"a2a_gemm {M}x{N}x{K} it{it}")
        U.barrier_all_host()
        ctx.finalize()


def case_gemm_rs():
    from triton_dist.ops.gemm_rs import create_gemm_rs_context, gemm_rs
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    shapes = [(256 * W, 512, 512), (128 * W, 256, 1024), (512 * W, 1024, 256), (128 * W, 264, 136), (40 * W, 256, 128),
              (256 * W, 5120, 2048), (256 * W, 2560, 4096)] if big else [(8 * W, 16, 24), (3 * W, 8, 8)]
    dtype = torch.bfloat16 if big else torch.float32
    for si, (M, N, K) in enumerate(shapes):
        # every other shape runs the fp32 ring (partial sums travel in fp32: one rounding at the owner)
        ctx = create_gemm_rs_context(M, N, output_dtype=dtype, fp32_ring=bool(big and si % 2 == 1))
        for it in range(5):
            A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
            Wt = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
            straggler = (it % W, 3_000_000) if (big and it in (1, 3)) else None
            C = gemm_rs(A, Wt.t(), ctx, straggler_option=straggler)
            full = (A.float() @ Wt.float().t())
            if dev.type == "cuda":
                ref = torch.empty(M // W, N, device=dev, dtype=torch.float32)
                dist.reduce_scatter_tensor(ref, full, group=U.get_triton_dist_world())
            else:   # gloo has no reduce_scatter_tensor
                dist.all_reduce(full, group=U.get_triton_dist_world())
                ref = full[me * (M // W):(me + 1) * (M // W)]
            _assert_close(C, ref, (0.25 if ctx.fp32_ring else 1.0) if big else 1e-3, 3e-2 if big else 1e-4,
                          f"gemm_rs {M}x{N}x{K} it{it} fp32_ring={ctx.fp32_ring}")
        U.barrier_all_host()
        ctx.finalize()


def case_gemm_ar():
    """GEMM + AllReduce: two-kernel path and the single fused kernel (gemm_allreduce.py:565-604,669-731)."""
    from triton_dist.ops.gemm_ar import (create_gemm_ar_context, create_ll_gemm_ar_context, gemm_allreduce_op,
                                         low_latency_gemm_allreduce_op)
    from triton_dist.ops.gemm import GemmConfig
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    shapes = [(128, 4096, 512), (16, 1024, 1024), (256, 2048, 256), (200, 520, 264), (512, 1024, 512)] if big else [(8, 16, 24)]
    for (M, N, K) in shapes:
        ctx = create_ll_gemm_ar_context(None, me, W, W, max_M=M, N=N, dtype=dtype)
        cfgs = [None]
        if big:
            cfgs += [GemmConfig(128, 2, 8, False, 0, 16), GemmConfig(256, 1, 8, False, 0, 8)]
        for it in range(6):
            A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
            Wt = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
            ref = A.float() @ Wt.float().t()
            dist.all_reduce(ref, group=U.get_triton_dist_world())
            straggler = (it % W, 2_000_000) if (big and it in (1, 4)) else None
            out = low_latency_gemm_allreduce_op(ctx, A, Wt, gemm_config=cfgs[it % len(cfgs)], straggler_option=straggler)
            _assert_close(out, ref, 1.0 if big else 1e-3, 3e-2 if big else 1e-4, f"ll_gemm_ar {M}x{N}x{K} it{it}")
            out2 = gemm_allreduce_op(ctx, A, Wt)
            _assert_close(out2, ref, 1.0 if big else 1e-3, 3e-2 if big else 1e-4, f"gemm_ar {M}x{N}x{K} it{it}")
        U.barrier_all_host()
        ctx.finalize()


def case_a2a():
    """Device-split all-to-all family (low_latency_all_to_all.py / all_to_all_vdev_2d_offset.py): packed splits and the offset
    variant (rows of every (rank, expert) segment start at arbitrary offsets of the send buffer) vs a gathered golden."""
    from triton_dist.ops.all_to_all import all_to_all_vdev_2d, create_all_to_all_context
    from triton_dist.ops.compat import all_to_all_vdev_2d_offset
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    g_, Hd, max_m = 2, 64, 48
    ctx = create_all_to_all_context(max_m, Hd, experts_per_rank=g_, dtype=dtype)
    grp = U.get_triton_dist_world()
    for it in range(4):
        gen = torch.Generator().manual_seed(31 + it)                      # every rank draws ALL ranks' splits (same seed)
        all_sp = torch.randint(0, 7, (W, W * g_), generator=gen)
        sp = all_sp[me].to(torch.int32).to(dev)
        n = int(all_sp[me].sum())
        x = (torch.randn(n, Hd, generator=torch.Generator().manual_seed(100 * it + me)) * 0.5).to(dtype).to(dev)
        out, out_sp = all_to_all_vdev_2d(ctx, x, sp)
        # golden: rank s's rows for me are its segments [me*g_, (me+1)*g_)
        exp = []
        for s_ in range(W):
            xs = (torch.randn(int(all_sp[s_].sum()), Hd, generator=torch.Generator().manual_seed(100 * it + s_)) * 0.5).to(dtype)
            c = torch.cumsum(all_sp[s_], 0)
            lo, hi = int(c[me * g_] - all_sp[s_][me * g_]), int(c[(me + 1) * g_ - 1])
            exp.append(xs[lo:hi])
        exp = torch.cat(exp)
        assert torch.equal(out.cpu().float(), exp.float()), f"a2a it{it}"
        assert torch.equal(out_sp.cpu(), all_sp[:, me * g_:(me + 1) * g_].to(torch.int32))
        # offset variant: the same segments scattered over a larger buffer with gaps
        gap = 3
        offs = (torch.cumsum(all_sp[me] + gap, 0) - all_sp[me] - gap + 1).to(torch.int32)
        big_buf = torch.zeros(n + gap * W * g_ + 4, Hd, dtype=dtype, device=dev)
        c = torch.cumsum(all_sp[me], 0) - all_sp[me]
        for j in range(W * g_):
            big_buf[int(offs[j]):int(offs[j]) + int(all_sp[me][j])] = x[int(c[j]):int(c[j]) + int(all_sp[me][j])]
        out2, _ = all_to_all_vdev_2d_offset(ctx, big_buf, sp, offs.to(dev), g_)
        assert torch.equal(out2.cpu().float(), exp.float()), f"a2a offset it{it}"
    U.barrier_all_host()
    ctx.finalize()


def case_gemm_q8():
    """Quantised fused ops: int8 x scale GEMM + AllReduce (reference gemm_allreduce.py:383-447) and int8 / per-tensor fp8 gemm_rs
    (test_gemm_rs.py:130-145) vs fp32 math on the dequantised operands + NCCL."""
    from triton_dist.ops.gemm_ar import create_gemm_ar_context, gemm_allreduce_op
    from triton_dist.ops.gemm_rs import create_gemm_rs_context, gemm_rs
    dev = U.current_device()
    if dev.type != "cuda":
        return
    W, me = U.world_size(), U.rank()
    grp = U.get_triton_dist_world()
    for (M, N, K) in [(128 * W, 512, 256), (256 * W, 1280, 384)]:
        actx = create_gemm_ar_context(None, me, W, W, M, N, torch.bfloat16)
        rctx = create_gemm_rs_context(M, N, output_dtype=torch.bfloat16)
        for it, kind in enumerate(["int8", "fp8", "int8"]):
            if kind == "int8":
                a = torch.randint(-127, 128, (M, K), device=dev, dtype=torch.int8)
                b = torch.randint(-127, 128, (N, K), device=dev, dtype=torch.int8)
            else:
                a = (torch.randn(M, K, device=dev) * 0.5).to(torch.float8_e4m3fn)
                b = (torch.randn(N, K, device=dev) * 0.5).to(torch.float8_e4m3fn)
            sa = torch.rand(M, device=dev) * 0.02 + 0.001
            sb = torch.rand(N, device=dev) * 0.02 + 0.001
            full = (a.float() @ b.float().t()) * sa[:, None] * sb[None, :]
            ref_ar = full.clone()
            dist.all_reduce(ref_ar, group=grp)
            out = gemm_allreduce_op(actx, a, b, As=sa, Bs=sb)
            tol = 2e-2 * ref_ar.abs().max().item() + 1e-2
            _assert_close(out, ref_ar, tol, 3e-2, f"gemm_ar {kind} {M}x{N}x{K} it{it}")
            ref_rs = torch.empty(M // W, N, device=dev, dtype=torch.float32)
            dist.reduce_scatter_tensor(ref_rs, full, group=grp)
            out2 = gemm_rs(a, b.t(), rctx, scale_a=sa, scale_b=sb)
            _assert_close(out2, ref_rs, tol, 3e-2, f"gemm_rs {kind} {M}x{N}x{K} it{it}")
        U.barrier_all_host()
        actx.finalize(); rctx.finalize()


def case_gemm_a2a():
    """GEMM with the all-to-all of its output columns fused into the epilogue (ulysses_sp_infer_gemm_a2a.py:143-289)."""
    from triton_dist.ops.gemm_a2a import create_gemm_a2a_context, gemm_all_to_all
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    shapes = [(256, 256, 512), (100, 128, 264), (512, 384, 1024), (32, 64, 128)] if big else [(6, 8, 16)]
    for (M, c, K) in shapes:
        ctx = create_gemm_a2a_context(M, c, dtype)
        for it in range(5):
            X = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
            Wt = (torch.randn(W * c, K, device=dev) * 0.5).to(dtype)
            if it in (1, 3) and big and me == it % W:
                torch.cuda._sleep(2_000_000)
            out = gemm_all_to_all(ctx, X, Wt)
            # golden: every rank's full product, gathered; my columns of each
            y = (X.float() @ Wt.float().t())
            ally = torch.empty(W * M * W * c, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(ally, y.reshape(-1).contiguous(), group=U.get_triton_dist_world())
            ref = ally.view(W, M, W * c)[:, :, me * c:(me + 1) * c].reshape(W * M, c)
            _assert_close(out, ref, 0.5 if big else 1e-3, 2e-2 if big else 1e-4, f"gemm_a2a {M}x{c}x{K} it{it}")
        U.barrier_all_host()
        ctx.finalize()


def case_gemm_a2a_q8():
    """Quantised GEMM + all-to-all (Ulysses inference flavour): int8 x int8 with per-row input scales and per-channel weight scales,
    dequantised in the epilogue, bf16 on the wire (reference: ulysses_sp_infer_gemm_a2a.py:143-260).  GPU: two-kernel default and the
    fused one-kernel variant; emulation: the same flag protocol as the 16-bit op."""
    from triton_dist.ops.compat import ulysses_sp_infer_gemm_a2a_op
    from triton_dist.ops.gemm_a2a import create_gemm_a2a_context, gemm_all_to_all
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    M, K, c = 128, 256, 64
    ctx = create_gemm_a2a_context(M, c, torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    w8 = torch.randint(-8, 8, (W * c, K), generator=g, dtype=torch.int8).to(dev)            # same weight on every rank
    sb = (torch.rand(W * c, generator=g) * 0.05 + 0.01).to(dev)
    variants = (None, True) if dev.type == "cuda" else (None,)
    for it in range(3):
        for fused in variants:
            gx = torch.Generator().manual_seed(100 * it + me)
            x8 = torch.randint(-8, 8, (M, K), generator=gx, dtype=torch.int8).to(dev)
            sa = (torch.rand(M, generator=gx) * 0.1 + 0.02).to(dev)
            if fused is None:
                out = ulysses_sp_infer_gemm_a2a_op(ctx, x8, w8, input_scale=sa, weight_scale=sb)
            else:
                out = gemm_all_to_all(ctx, x8, w8, scale_a=sa, scale_b=sb, fused=True)
            # reference: every rank's dequantised product, my column block of each
            mine = ((x8.float() @ w8.float().t()) * sa[:, None] * sb[None, :]).to(torch.bfloat16)
            full = [torch.empty_like(mine) for _ in range(W)]
            dist.all_gather(full, mine, group=U.get_triton_dist_world())
            ref = torch.cat([f[:, me * c:(me + 1) * c] for f in full], 0)
            _assert_close(out, ref, 2e-2, 2e-2, f"gemm_a2a_q8 it {it} fused {fused}")
    U.barrier_all_host()
    ctx.finalize()


def case_moe():
    """ag_group_gemm + run_moe_reduce_rs vs the masked-matmul golden (reference: test_ag_moe.py, test_moe_reduce_rs.py)."""
    from triton_dist.ops import moe as M
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    T, H, I, E, topk = (256 * W, 512, 1024, 8, 2) if big else (8 * W, 16, 32, 4, 2)
    ag = M.create_ag_group_gemm_context(T, I // W, H, E, topk, dtype)
    rs = M.create_moe_rs_context(me, W, W, T * topk, H, E, topk, dtype)
    grp = U.get_triton_dist_world()
    for it in range(3):
        g = torch.Generator(device="cpu").manual_seed(100 + it)          # same routing on all ranks
        ids = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32).to(dev)
        wts = torch.softmax(torch.randn(T, topk, generator=g), -1).to(dev)
        x = (torch.randn(T // W, H, device=dev) * 0.5).to(dtype)
        w_up = (torch.randn(E, I // W, H, device=dev) * 0.2).to(dtype)
        w_dn = (torch.randn(E, H, I // W, device=dev) * 0.2).to(dtype)
        c = M.ag_group_gemm(x, w_up, ag, ids)
        xf = torch.empty(T * H, device=dev, dtype=dtype)
        dist.all_gather_into_tensor(xf, x.view(-1), group=grp)
        xf = xf.view(T, H)
        ref = torch.stack([xf[t].float() @ w_up[int(ids[t, j])].float().t() for t in range(T) for j in range(topk)])
        _assert_close(c, ref, 0.3 if big else 1e-3, 3e-2 if big else 1e-4, f"ag_group_gemm it{it}")
        h = (torch.randn(T * topk, I // W, device=dev) * 0.5).to(dtype)
        out = M.run_moe_reduce_rs(h, w_dn, ids, wts, rs)
        gold = M.moe_reduce_rs_torch(h, w_dn.transpose(1, 2), ids, wts, grp, W, me)
        _assert_close(out, gold, 0.5 if big else 1e-3, 3e-2 if big else 1e-4, f"moe_reduce_rs it{it}")
    U.barrier_all_host()
    ag.finalize(); rs.finalize()


def case_moe_rs():
    """Single-kernel grouped GEMM + weighted top-k reduce + ReduceScatter / AllReduce (mode kMoeRS) vs the masked-matmul golden;
    several shapes (256- and 128-wide n tiles, ragged last n tile), >= 4 calls per context (parity double buffering), a straggler."""
    from triton_dist.ops import moe as M
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    grp = U.get_triton_dist_world()
    shapes = [(256 * W, 512, 1024, 8, 2), (128 * W, 384, 512, 4, 2), (64 * W, 1032, 256, 8, 4), (1024 * W, 1280, 512, 8, 2)] if big else [(8 * W, 16, 32, 4, 2)]
    for (T, H, I, E, topk) in shapes:
        rs = M.create_moe_rs_context(me, W, W, T * topk, H, E, topk, dtype)
        for it in range(5):
            g = torch.Generator(device="cpu").manual_seed(7 + it)          # same routing on all ranks
            ids = torch.rand(T, E, generator=g).topk(topk, dim=1).indices.to(torch.int32).to(dev)
            wts = torch.softmax(torch.randn(T, topk, generator=g), -1).to(dev)
            h = (torch.randn(T * topk, I // W, device=dev) * 0.5).to(dtype)
            w_dn = (torch.randn(E, H, I // W, device=dev) * 0.2).to(dtype)
            if big and it == 2:
                torch.cuda._sleep(3_000_000 * (1 + me))
            gold_full = M.moe_reduce_rs_torch(h, w_dn.transpose(1, 2), ids, wts, grp, 1, 0).float()      # my partial [T, H] (fp32)
            dist.all_reduce(gold_full, group=grp)
            if it % 2 == 0:
                out = M.run_moe_reduce_rs(h, w_dn, ids, wts, rs)
                _assert_close(out, gold_full[me * (T // W):(me + 1) * (T // W)], 0.5 if big else 1e-3, 3e-2 if big else 1e-4,
                              f"moe_reduce_rs T{T} H{H} it{it}")
            else:
                out = M.run_moe_reduce_ar(h, w_dn, ids, wts, rs)
                _assert_close(out, gold_full, 0.5 if big else 1e-3, 3e-2 if big else 1e-4, f"moe_reduce_ar T{T} H{H} it{it}")
        U.barrier_all_host()
        rs.finalize()


def case_ep_mega():
    """Mega-EP (dispatch || grouped GEMM, grouped GEMM || combine: csrc/gemm_sm100.cuh kEPD / kEPC) vs the fp32 golden computed
    with all experts' weights; unbalanced routing (some experts empty), several calls per context (parity), a straggler."""
    from triton_dist.ops import ep_mega as EM
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    bf = torch.bfloat16 if big else torch.float32
    shapes = [(512, 512, 256, 4, 2), (300, 1024, 512, 2, 4), (1024, 256, 128, 8, 3)] if big else [(24, 16, 8, 2, 2), (10, 8, 8, 3, 3)]
    for (T, H, I, epr, topk) in shapes:
        E = epr * W
        ctx = EM.create_ep_mega_context(T, H, topk, E, bf, capacity_factor=3.0)
        g = torch.Generator(device="cpu").manual_seed(11)
        w_gu_all = (torch.randn(E, 2 * I, H, generator=g) * 0.05).to(bf).to(dev)
        w_dn_all = (torch.randn(E, H, I, generator=g) * 0.05).to(bf).to(dev)
        w_gu, w_dn = w_gu_all[me * epr:(me + 1) * epr].contiguous(), w_dn_all[me * epr:(me + 1) * epr].contiguous()
        for it in range(4):
            x = (torch.randn(T, H, device=dev) * 0.5).to(bf)
            logits = torch.randn(T, E, device=dev)
            if it >= 2:
                logits[:, E // 2:] -= 4.0          # unbalanced: the upper half of the experts is (almost) never chosen
            ids = logits.topk(topk, dim=1).indices.to(torch.int32)
            wts = torch.softmax(torch.randn(T, topk, device=dev), -1)
            if it == 1 and big:
                torch.cuda._sleep(2_000_000 * (1 + me))
            out = EM.mega_ep_moe(ctx, x, ids, wts, w_gu, w_dn)
            ref = EM.mega_ep_moe_reference(x, ids, wts, w_gu_all, w_dn_all)
            _assert_close(out, ref, 0.05 if big else 1e-4, 5e-2 if big else 1e-4, f"ep_mega T{T} H{H} it{it}")
        if not big:
            U.barrier_all_host()
            ctx.finalize()
            continue
        # training path: forward + backward through the same kernels vs torch autograd on the dense formulation
        from triton_dist.function.nvidia import mega_ep_moe_autograd
        for it in range(2):
            x = (torch.randn(T, H, device=dev) * 0.5).to(bf).requires_grad_(True)
            ids = torch.randn(T, E, device=dev).topk(topk, dim=1).indices.to(torch.int32)
            wts = torch.softmax(torch.randn(T, topk, device=dev), -1).requires_grad_(True)
            wg, wd = w_gu.clone().requires_grad_(True), w_dn.clone().requires_grad_(True)
            g_out = (torch.randn(T, H, device=dev) * 0.1).to(bf)
            out = mega_ep_moe_autograd(ctx, x, ids, wts, wg, wd)
            out.backward(g_out)
            # dense golden (fp32 math, bf16 rounding where the kernels round), gradients of ALL experts summed over ranks
            xr = x.detach().float().requires_grad_(True)
            wr = wts.detach().float().requires_grad_(True)
            ga, da = w_gu_all.float().requires_grad_(True), w_dn_all.float().requires_grad_(True)
            o = torch.zeros(T, H, device=dev)
            for k in range(topk):
                e = ids[:, k].long()
                hh = torch.einsum("th,tih->ti", xr, ga[e])
                a = torch.nn.functional.silu(hh[:, :I]) * hh[:, I:]
                o = o + wr[:, k:k + 1] * torch.einsum("ti,thi->th", a, da[e])
            o.backward(g_out.float())
            dist.all_reduce(ga.grad, group=U.get_triton_dist_world()); dist.all_reduce(da.grad, group=U.get_triton_dist_world())
            _assert_close(out.detach(), o.detach(), 0.05, 5e-2, f"ep_mega autograd fwd it{it}")
            _assert_close(x.grad, xr.grad, 0.05, 6e-2, f"ep_mega dX it{it}")
            _assert_close(wts.grad, wr.grad, 0.08, 6e-2, f"ep_mega d(routing w) it{it}")
            sl = slice(me * epr, (me + 1) * epr)
            scale_g = max(1.0, ga.grad[sl].abs().max().item())
            _assert_close(wg.grad / scale_g, ga.grad[sl] / scale_g, 0.03, 6e-2, f"ep_mega dW_gate_up it{it}")
            scale_d = max(1.0, da.grad[sl].abs().max().item())
            _assert_close(wd.grad / scale_d, da.grad[sl] / scale_d, 0.03, 6e-2, f"ep_mega dW_down it{it}")
        U.barrier_all_host()
        ctx.finalize()


def case_moe_staged():
    """The multi-kernel MoE path (all-gather kernel -> gather_rows -> grouped GEMM -> scatter_rows); ``case_moe`` runs the
    default single-kernel path on GPUs (AllGather + grouped GEMM with a TMA tile::gather4 producer waiting on arrival flags)."""
    os.environ["TD_MOE_AG_FUSED"] = "0"
    os.environ["TD_MOE_TMA_GATHER"] = "0"
    try:
        case_moe()
    finally:
        os.environ.pop("TD_MOE_AG_FUSED", None)
        os.environ.pop("TD_MOE_TMA_GATHER", None)


def case_moe_fused():
    case_moe()


def case_tp_e2e():
    """TP inference demo: every backend must reproduce the torch (NCCL) backend's greedy tokens (test_tp_e2e.py --check)."""
    from triton_dist.models import Engine, ModelConfig
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    for name in ("tiny-dense", "tiny-moe"):
        cfg = ModelConfig(model_name=name, max_length=64, dtype=torch.bfloat16 if big else torch.float32, rank=me, world_size=W)
        eng = Engine(cfg, temperature=0.0)
        g = torch.Generator().manual_seed(7)
        ids = torch.randint(0, 1000, (2 * W, 6), generator=g)
        ref = eng.serve(ids, 5, backend="torch", use_cuda_graph=False)
        backends = ("triton_dist", "triton_dist_AR", "triton_dist_gemm_ar") if name == "tiny-dense" else ("triton_dist", "triton_dist_AR")
        for be in backends:
            out = eng.serve(ids, 5, backend=be, use_cuda_graph=big)
            agree = (out == ref).float().mean().item()
            assert agree >= (0.7 if big else 1.0), (name, be, agree, out.tolist(), ref.tolist())
            eng.model.finalize()
        U.barrier_all_host()


def case_engine_mega():
    """``Engine.serve(backend="mega")``: prefill per-op, then every decode step is one persistent-kernel launch on the engine's own model and
    KV cache; greedy tokens must match the torch backend.  MoE models are rejected."""
    from triton_dist.models import Engine, ModelConfig
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    cfg = ModelConfig(model_name="tiny-dense", max_length=64, dtype=torch.bfloat16 if big else torch.float32, rank=me, world_size=W)
    eng = Engine(cfg, temperature=0.0)
    ids = torch.randint(0, 1000, (3, 5), generator=torch.Generator().manual_seed(9))
    ref = eng.serve(ids, 6, backend="torch", use_cuda_graph=False)
    for graph in ((False, True) if big else (False,)):
        out = eng.serve(ids, 6, backend="mega", use_cuda_graph=graph)
        agree = (out == ref).float().mean().item()
        assert agree >= (0.7 if big else 1.0), (graph, agree, out.tolist(), ref.tolist())
    eng.finalize()
    moe = Engine(ModelConfig(model_name="tiny-moe", max_length=64, dtype=cfg.dtype, rank=me, world_size=W), temperature=0.0)
    try:
        moe.serve(ids, 2, backend="mega", use_cuda_graph=False)
        raise AssertionError("the mega backend must reject MoE models")
    except ValueError:
        pass
    U.barrier_all_host()
    moe.finalize()


def case_ep_ll():
    """EP low-latency dispatch + combine vs a gathered golden (reference: test_ep_ll_a2a.py)."""
    from triton_dist.ops import ep_a2a as EP
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    T, H, topk, E = (96, 1024, 4, 8 * W) if big else (6, 128, 2, 2 * W)
    epr = E // W
    grp = U.get_triton_dist_world()
    variants = [False, True] if big else [False]
    for fp8 in variants:
        ctx = EP.create_ep_ll_a2a_ctx(T, H, topk, E, online_quant_fp8=fp8, dtype=torch.bfloat16)
        for it in range(4):
            g = torch.Generator().manual_seed(1000 * it + me)
            x = (torch.randn(T, H, generator=g) * 0.5).to(torch.bfloat16).to(dev)
            idx = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32).to(dev)
            if it == 2:
                idx[0, 0] = -1                                   # an unrouted slot
            wts = torch.softmax(torch.randn(T, topk, generator=g), -1).to(dev)
            rx, rs, cnt, meta = EP.ep_ll_dispatch(ctx, x, idx)
            # golden: gather everybody's tokens / routing
            xs = torch.empty(W * T * H, dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(xs, x.view(-1), group=grp)
            ids = torch.empty(W * T * topk, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(ids, idx.view(-1), group=grp)
            xs, ids = xs.view(W, T, H), ids.view(W, T, topk)
            cnts, starts = meta.counts_and_starts()
            dense = EP.dequant_fp8(rx, rs) if fp8 else rx
            for le in range(epr):
                e = me * epr + le
                total = 0
                for src in range(W):
                    want = (ids[src] == e).nonzero()                       # [n, 2] (token, k)
                    c, s = int(cnts[le, src]), int(starts[le, src])
                    assert c == want.shape[0], (le, src, c, want.shape[0])
                    total += c
                    got_flat = meta.recv_token_source_indices[le, s:s + c].long()
                    assert sorted(got_flat.tolist()) == sorted((want[:, 0] * topk + want[:, 1]).tolist())
                    ref_rows = xs[src][got_flat // topk].float()
                    _assert_close(dense[le, s:s + c], ref_rows, 0.08 if fp8 else 0.0, 0.08 if fp8 else 0.0, f"dispatch payload fp8={fp8}")
                assert int(cnt[le]) == total
            # experts = identity * (1 + local expert id): combine must return sum_k w * (1 + le) * x
            y = torch.zeros((epr, W * T, H), dtype=torch.bfloat16, device=dev)
            for le in range(epr):
                n = int(cnt[le])
                y[le, :n] = (dense[le, :n].float() * (1 + me * epr + le)).to(torch.bfloat16)
            out = EP.ep_ll_combine(ctx, y, idx, wts, meta)
            scale = ((idx.clamp(min=0).float() + 1) * wts * (idx >= 0)).sum(-1, keepdim=True)
            _assert_close(out, x.float() * scale, 0.35 if fp8 else 0.1, 0.1 if fp8 else 3e-2, f"combine fp8={fp8} it{it}")
        U.barrier_all_host()
        ctx.finalize()


def case_ep_normal():
    """Throughput-mode EP (token saving): dispatch -> expert FFN through the index lists -> local pre-reduce + combine, vs a golden
    that evaluates every expert on every token (reference: test_ep_a2a.py / ep_a2a_intra_node.py)."""
    from triton_dist.ops import ep_normal as EN
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    T, H, I, topk, epr = (200, 512, 256, 4, 4) if big else (7, 16, 8, 3, 2)
    E = epr * W
    ctx = EN.create_ep_normal_ctx(T, H, topk, E, dtype)
    gw = torch.Generator().manual_seed(7)                                  # all ranks draw the same expert weights
    w_gu = (torch.randn(E, 2 * I, H, generator=gw) * 0.2).to(dtype)
    w_dn = (torch.randn(E, H, I, generator=gw) * 0.2).to(dtype)
    my_gu, my_dn = w_gu[me * epr:(me + 1) * epr].contiguous().to(dev), w_dn[me * epr:(me + 1) * epr].contiguous().to(dev)
    for it in range(3):
        g = torch.Generator().manual_seed(100 * it + me)
        Tn = T if it != 1 else T - 3                                         # a shorter batch re-uses the same buffers
        x = (torch.randn(Tn, H, generator=g) * 0.5).to(dtype)
        idx = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(Tn)]).to(torch.int32)
        if it == 2:
            idx[0, 0] = -1
        wts = torch.softmax(torch.randn(Tn, topk, generator=g), -1)
        h = EN.ep_dispatch_normal(ctx, x.to(dev), idx.to(dev), wts.to(dev))
        # token saving: rows received == distinct (token, rank) pairs of every source
        allidx = [None] * W
        dist.all_gather_object(allidx, idx, group=U._gloo_group())
        for src in range(W):
            want_rows = sum(len({int(e) // epr for e in row if int(e) >= 0} & {me}) for row in allidx[src])
            want_pairs = sum(sum(1 for e in row if int(e) >= 0 and int(e) // epr == me) for row in allidx[src])
            assert int(h.rcnt[src, 0]) == want_rows and int(h.rcnt[src, 1]) == want_pairs, (src, h.rcnt[src].tolist(), want_rows, want_pairs)
        y = EN.ep_expert_ffn_normal(ctx, h, my_gu, my_dn)
        out = EN.ep_combine_normal(ctx, y, h, idx.to(dev))
        ref = torch.zeros(Tn, H)
        for t in range(Tn):
            for k in range(topk):
                e = int(idx[t, k])
                if e < 0:
                    continue
                hh = x[t].float() @ w_gu[e].float().t()
                act = torch.nn.functional.silu(hh[:I]) * hh[I:]
                ref[t] += float(wts[t, k]) * (act.to(dtype).float() @ w_dn[e].float().t())
        _assert_close(out, ref, 0.3 if big else 1e-4, 5e-2 if big else 1e-4, f"ep_normal it{it}")
    U.barrier_all_host()
    ctx.finalize()
    if not big:
        # the Mega-EP style op object (lazy sizing -> materialize -> dispatch+GEMM half -> GEMM+combine half)
        from triton_dist.parallel.ep import EPConfig, EpAll2AllFusedOp
        op = EpAll2AllFusedOp(EPConfig(T, H, topk, E, me, W, False, dtype))
        assert op.get_nvshmem_size() > 0 and op.ctx is None
        op.materialize()
        g = torch.Generator().manual_seed(555 + me)
        x = (torch.randn(T, H, generator=g) * 0.5).to(dtype)
        idx = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32)
        wts = torch.softmax(torch.randn(T, topk, generator=g), -1)
        assert int(op.preprocess(idx).sum()) == T * topk
        act, handle = op.mega_dispatch_group_gemm(x, idx, wts, my_gu)
        out = op.mega_group_gemm_combine(act, handle, my_dn)
        ref = torch.zeros(T, H)
        for t in range(T):
            for k in range(topk):
                e = int(idx[t, k])
                hh = x[t].float() @ w_gu[e].float().t()
                ref[t] += float(wts[t, k]) * ((torch.nn.functional.silu(hh[:I]) * hh[I:]).to(dtype).float() @ w_dn[e].float().t())
        _assert_close(out, ref, 1e-4, 1e-4, "EpAll2AllFusedOp")
        U.barrier_all_host()
        op.finalize()


def case_sp_pp():
    """Ulysses a2a round trip, SP flash-decode (KV sharded over ranks), AG-KV context-parallel attention, PP send/recv."""
    import math
    from triton_dist.parallel.sp import (SpGQAFlashDecodeAttention, UlyssesSPAllToAllLayer, create_sp_ag_attention_context_intra_node,
                                         fused_sp_ag_attn_intra_node, zigzag_positions)
    from triton_dist.parallel.pp import PPCommLayer
    from triton_dist.ops.flash_decode import _decode_reference
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    grp = U.get_triton_dist_world()
    g = torch.Generator().manual_seed(5)
    # ---- Ulysses ----
    S_l, H, D = 8, 2 * W, 128
    uly = UlyssesSPAllToAllLayer(S_l, H, D, dtype, me, W)
    full = torch.randn(S_l * W, H, D, generator=g).to(dtype).to(dev)
    x = full[me * S_l:(me + 1) * S_l].contiguous()
    y = uly.pre_attn_a2a(x)
    assert torch.equal(y.cpu(), full[:, me * (H // W):(me + 1) * (H // W)].cpu())
    z = uly.post_attn_a2a(y)
    assert torch.equal(z.cpu(), x.cpu())
    uly.finalize()
    # ---- SP flash decode ----
    B, Hq, Hkv, L_l = 2, 4, 1, 40
    q = torch.randn(B, Hq, D, generator=g).to(dtype).to(dev)
    kf = torch.randn(B, L_l * W, Hkv, D, generator=g).to(dtype).to(dev)
    vf = torch.randn(B, L_l * W, Hkv, D, generator=g).to(dtype).to(dev)
    lens_full = torch.tensor([L_l * W, L_l * W - 3], dtype=torch.int32, device=dev)
    local_lens = (lens_full - me * L_l).clamp(0, L_l).to(torch.int32)
    sp = SpGQAFlashDecodeAttention(me, W, Hq, Hkv, D, max_batch=B)
    out = sp(q, kf[:, me * L_l:(me + 1) * L_l].contiguous(), vf[:, me * L_l:(me + 1) * L_l].contiguous(), local_lens)
    ref, _ = _decode_reference(q, kf, vf, lens_full, 1 / math.sqrt(D))
    _assert_close(out, ref, 3e-2 if big else 1e-4, 3e-2 if big else 1e-4, "sp flash decode")
    sp.finalize()
    # ---- AG-KV context parallel attention (zig-zag) ----
    S = 16 * W
    qf = torch.randn(S, Hq, D, generator=g).to(dtype).to(dev)
    kf2 = torch.randn(S, Hkv, D, generator=g).to(dtype).to(dev)
    vf2 = torch.randn(S, Hkv, D, generator=g).to(dtype).to(dev)
    pos = zigzag_positions(S, W, me, dev) if W > 1 else torch.arange(S, device=dev)
    ctx = create_sp_ag_attention_context_intra_node(S // W, Hkv, D, dtype)
    o = fused_sp_ag_attn_intra_node(ctx, qf[pos].contiguous(), kf2[pos].contiguous(), vf2[pos].contiguous(), is_causal=True)
    kk, vv = kf2.float().repeat_interleave(Hq // Hkv, 1), vf2.float().repeat_interleave(Hq // Hkv, 1)
    sc = torch.einsum("shd,lhd->hsl", qf.float(), kk) / math.sqrt(D)
    sc = sc.masked_fill(~(torch.arange(S, device=dev)[None, :] <= torch.arange(S, device=dev)[:, None])[None], float("-inf"))
    full_o = torch.einsum("hsl,lhd->shd", torch.softmax(sc, -1), vv)
    _assert_close(o, full_o[pos], 3e-2 if big else 1e-4, 3e-2 if big else 1e-4, "sp ag attention")
    ctx.finalize()
    if big:     # long enough for the tcgen05 flash kernel (zig-zag chunks are multiples of the 128-query tile)
        S = 512 * W
        qf = torch.randn(S, Hq, D, generator=g).to(dtype).to(dev)
        kf2 = torch.randn(S, Hkv, D, generator=g).to(dtype).to(dev)
        vf2 = torch.randn(S, Hkv, D, generator=g).to(dtype).to(dev)
        pos = zigzag_positions(S, W, me, dev) if W > 1 else torch.arange(S, device=dev)
        ctx = create_sp_ag_attention_context_intra_node(S // W, Hkv, D, dtype)
        o = fused_sp_ag_attn_intra_node(ctx, qf[pos].contiguous(), kf2[pos].contiguous(), vf2[pos].contiguous(), is_causal=True)
        kk, vv = kf2.float().repeat_interleave(Hq // Hkv, 1), vf2.float().repeat_interleave(Hq // Hkv, 1)
        sc = torch.einsum("shd,lhd->hsl", qf.float(), kk) / math.sqrt(D)
        sc = sc.masked_fill(~(torch.arange(S, device=dev)[None, :] <= torch.arange(S, device=dev)[:, None])[None], float("-inf"))
        full_o = torch.einsum("hsl,lhd->shd", torch.softmax(sc, -1), vv)
        _assert_close(o, full_o[pos], 3e-2, 3e-2, "sp ag attention (tcgen05 flash, zig-zag)")
        # the same attention with the KV all-gather overlapped with the local-chunk flash call + LSE merge of the partials
        from triton_dist.parallel.sp import fused_sp_ag_attn_overlapped
        for zz in (True, False):
            pp = pos if zz else torch.arange(me * (S // W), (me + 1) * (S // W), device=dev)
            o2 = fused_sp_ag_attn_overlapped(ctx, qf[pp].contiguous(), kf2[pp].contiguous(), vf2[pp].contiguous(), is_causal=True, enable_zig_zag=zz)
            _assert_close(o2, full_o[pp], 3e-2, 3e-2, f"sp ag attention overlapped (zig-zag={zz})")
        ctx.finalize()
    # ---- PP send/recv ring ----
    for backend in (("triton_dist", "torch") if big else ("triton_dist",)):
        pp = PPCommLayer(1024, dtype, me, W, backend=backend, group=grp)
        for it in range(5):
            t = torch.full((4, 64), float(me * 10 + it), dtype=dtype, device=dev)
            if me % 2 == 0:
                pp.send(t, (me + 1) % W)
                got = pp.recv((4, 64), dtype, (me - 1) % W)
            else:
                got = pp.recv((4, 64), dtype, (me - 1) % W)
                pp.send(t, (me + 1) % W)
            assert torch.all(got.float() == ((me - 1) % W) * 10 + it), (backend, it, got[0, :3])
        U.barrier_all_host()
        pp.finalize()


def case_sp_varlen():
    """Context-parallel attention over a packed variable-length batch: every rank holds len_b / W tokens of each sequence (zig-zag
    inside each sequence), ONE gather of the packed KV, attention per sequence (reference: sp_ag_attention_intra_node.py:279-360)."""
    import math
    from triton_dist.parallel.sp import create_sp_ag_attention_context_intra_node, fused_sp_ag_attn_varlen, zigzag_positions
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    Hq, Hkv, D = 4, 2, 128
    unit = 256 * W if big else 4 * W            # sequence lengths: multiples of 2W (zig-zag) -- and of the 128-query tile on a GPU
    lens = [unit, 3 * unit, 2 * unit]
    g = torch.Generator().manual_seed(11)
    qs = [torch.randn(L, Hq, D, generator=g).to(dtype).to(dev) for L in lens]
    ks = [torch.randn(L, Hkv, D, generator=g).to(dtype).to(dev) for L in lens]
    vs = [torch.randn(L, Hkv, D, generator=g).to(dtype).to(dev) for L in lens]
    for zz in (True, False):
        def mine(L):
            return zigzag_positions(L, W, me, dev) if (zz and W > 1) else torch.arange(me * (L // W), (me + 1) * (L // W), device=dev)
        pos = [mine(L) for L in lens]
        q_sh = torch.cat([q[p] for q, p in zip(qs, pos)]).contiguous()
        k_sh = torch.cat([k[p] for k, p in zip(ks, pos)]).contiguous()
        v_sh = torch.cat([v[p] for v, p in zip(vs, pos)]).contiguous()
        cu = torch.tensor([0] + list(torch.tensor([L // W for L in lens]).cumsum(0)), dtype=torch.int32, device=dev)
        ctx = create_sp_ag_attention_context_intra_node(q_sh.shape[0], Hkv, D, dtype)
        for it in range(2):
            o = fused_sp_ag_attn_varlen(ctx, q_sh, k_sh, v_sh, cu, is_causal=True, enable_zig_zag=zz)
            ofs = 0
            for q, k, v, p, L in zip(qs, ks, vs, pos, lens):
                kk, vv = k.float().repeat_interleave(Hq // Hkv, 1), v.float().repeat_interleave(Hq // Hkv, 1)
                sc = torch.einsum("shd,lhd->hsl", q.float(), kk) / math.sqrt(D)
                sc = sc.masked_fill(~(torch.arange(L, device=dev)[None, :] <= torch.arange(L, device=dev)[:, None])[None], float("-inf"))
                full_o = torch.einsum("hsl,lhd->shd", torch.softmax(sc, -1), vv)
                _assert_close(o[ofs:ofs + L // W], full_o[p], 3e-2 if big else 1e-4, 3e-2 if big else 1e-4, f"sp varlen zz={zz} L={L}")
                ofs += L // W
        ctx.finalize()


def case_ep_moe():
    """EP_MoE layer (route -> dispatch -> grouped FFN -> combine) vs the gathered golden; autograd through dispatch/combine."""
    from triton_dist.parallel.ep import EP_MoE, TritonDistFusedEpMoeFunction
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    E, H, I, topk, T = 4 * W, 256 if big else 128, 128, 2, 64 if big else 6
    epr = E // W
    g = torch.Generator().manual_seed(11)
    router = (torch.randn(E, H, generator=g) * 0.5).to(dtype).to(dev)
    gall = (torch.randn(E, 2 * I, H, generator=g) * 0.1).to(dtype).to(dev)
    dall = (torch.randn(E, H, I, generator=g) * 0.1).to(dtype).to(dev)
    moe = EP_MoE(me, W, U.get_triton_dist_world())
    moe._init_parameters_from_shards(router, gall[me * epr:(me + 1) * epr].contiguous(), dall[me * epr:(me + 1) * epr].contiguous(), topk)
    moe._init_ctx(T)
    for it in range(3):
        x = (torch.randn(T, H, generator=torch.Generator().manual_seed(50 + it * W + me)) * 0.5).to(dtype).to(dev)
        out = moe.dist_triton_fwd(x)
        ref = moe.torch_fwd(x)
        _assert_close(out, ref, 5e-2 if big else 1e-3, 5e-2 if big else 1e-3, f"ep_moe it{it}")
    # gradient w.r.t. activations through dispatch/combine: compare with autograd of a dense single-process formulation
    if not big:
        x = torch.randn(T, H, generator=torch.Generator().manual_seed(99 + me)).to(dtype).to(dev).requires_grad_(True)
        ids, w = moe._route(x.detach())
        w = w.detach().clone().requires_grad_(True)
        y = TritonDistFusedEpMoeFunction.apply_moe(x, ids, w, moe.a2a, moe.w_gate_up, moe.w_down)
        (y * torch.linspace(0.5, 1.5, H)).sum().backward()
        xr = x.detach().clone().requires_grad_(True)
        wr = w.detach().clone().requires_grad_(True)
        acc = torch.zeros(T, H)
        for k in range(topk):
            for t in range(T):
                e = int(ids[t, k])
                h = gall[e].float() @ xr[t].float()
                h = torch.nn.functional.silu(h[:I]) * h[I:]
                acc[t] = acc[t] + wr[t, k] * (dall[e].float() @ h)
        (acc * torch.linspace(0.5, 1.5, H)).sum().backward()
        _assert_close(x.grad, xr.grad, 1e-3, 1e-3, "ep autograd dx")
        _assert_close(w.grad, wr.grad, 1e-3, 1e-3, "ep autograd d(routing weights)")
    U.barrier_all_host()
    moe.finalize()
    if not big or U.get_bool_env("TD_EP_NORMAL_GPU", False):     # GPU: the three stages are validated by case_ep_normal
        # the same layer on the throughput-mode (token saving) exchange
        moe._init_ctx(T, mode="normal")
        for it in range(2):
            x = (torch.randn(T, H, generator=torch.Generator().manual_seed(70 + it * W + me)) * 0.5).to(dtype).to(dev)
            _assert_close(moe.dist_triton_fwd(x), moe.torch_fwd(x), 5e-2 if big else 1e-3, 5e-2 if big else 1e-3, f"ep_moe normal it{it}")
        U.barrier_all_host()
        moe.finalize()


def case_lk():
    """Kernels written in the Python DSL (triton_dist.lk) on the symmetric heap: ring shift + push all-gather.  GPU backend: the
    generated CUDA is launched; emulation backend: the same Python source runs in the CPU interpreter (threads emulated), with
    symm_at / notify / wait on the shared-memory heap."""
    from triton_dist import lk
    from triton_dist.lk.kernels import simt as K
    W, me = U.world_size(), U.rank()
    n = 200
    dst = U.nvshmem_create_tensor((n,), torch.float32)
    out = U.nvshmem_create_tensor((W * n,), torch.float32)
    flags = U.nvshmem_create_tensor((16 + W,), torch.int32)
    flags.zero_()
    U.barrier_all_on_stream()
    ctx = lk.symm_ctx()
    gpu = dst.is_cuda

    def run(kernel, grid, *args):
        if gpu:
            kernel[grid](*args)
        else:
            kernel.interpret(grid, *args)

    for phase in range(1, 4):
        src = torch.arange(n, dtype=torch.float32, device=dst.device) + 1000.0 * me + phase
        run(K.ring_shift, 1, ctx, src, dst, flags[0:1], n, phase)
        prev = (me - 1 + W) % W
        _assert_close(dst, torch.arange(n, dtype=torch.float32) + 1000.0 * prev + phase, 0, 0, f"lk ring_shift phase {phase}")
        shard = torch.full((n,), float(me * 10 + phase), dtype=torch.float32, device=dst.device)
        run(K.allgather_push, W, ctx, shard, out, flags[16:], n, phase)
        want = torch.cat([torch.full((n,), float(r * 10 + phase)) for r in range(W)])
        _assert_close(out, want, 0, 0, f"lk allgather_push phase {phase}")
        U.barrier_all_on_stream()        # nobody overwrites dst / out of a rank that is still checking them
    for t in (flags, out, dst):
        U.nvshmem_free_tensor_sync(t)


SHMEM_TEST_SRC = r"""
#include "td/shmem.cuh"
using namespace td;
// every function family of the NVSHMEM-style header once: fcollect over the world team, broadcast inside the even team,
// put-with-signal around the ring (64-bit signals, CMP_GE wait), thread / warp scoped puts, team translation, barrier_all
__global__ void shmem_selftest(SymmCtx c, uint32_t* slots, uint32_t* epoch, float* fc_dst, float* bc_dst, float* ring_dst,
                               uint64_t* sig, int* misc, const float* src, int n, uint64_t phase) {
  shmem::Sync s{slots, epoch};
  const shmem::Team world = shmem::team_world(c);
  const shmem::Team even{0, 2, (c.world + 1) / 2};
  shmem::fcollect_block(c, world, s, fc_dst, src, (size_t)n);
  shmem::broadcast_block(c, even, s, bc_dst, src, (size_t)n, /*root=*/(int)(phase % even.size));
  const int nxt = (shmem::my_pe(c) + 1) % shmem::n_pes(c);
  shmem::putmem_signal_block(c, ring_dst, src, (size_t)n * sizeof(float), sig, phase, shmem::SIGNAL_SET, nxt);
  if (threadIdx.x == 0) {
    shmem::signal_wait_until(sig, shmem::CMP_GE, phase);
    misc[0] = shmem::team_translate_pe(world, c.rank, even);            // my index in the even team or -1
    misc[1] = shmem::team_my_pe(c, even);
    shmem::int_p(c, misc + 2, c.rank * 10 + (int)phase, nxt);           // thread-scope put of one int
  }
  if (threadIdx.x < 32) shmem::putmem_warp(c, ring_dst + n, src, 3 * sizeof(float) + 2, nxt);   // unaligned size: byte tail
  shmem::quiet();
  shmem::barrier_all_block(c, s);
}
extern "C" void launch_shmem_selftest(SymmCtx c, void* slots, void* epoch, void* fc, void* bc, void* ring, void* sig, void* misc,
                                      void* src, int n, unsigned long long phase, void* stream) {
  shmem_selftest<<<1, 128, 0, (cudaStream_t)stream>>>(c, (uint32_t*)slots, (uint32_t*)epoch, (float*)fc, (float*)bc, (float*)ring,
                                                      (uint64_t*)sig, (int*)misc, (const float*)src, n, phase);
}
"""


def case_shmem():
    """NVSHMEM-style API: the Python (host-initiated / stream-ordered) mirror on both backends, the device header csrc/td/shmem.cuh
    through a JIT kernel on the GPU backend (reference: test_nvshmem_api.py, test_team_split.py, test_ring_put.py)."""
    from triton_dist.language import shmem as S
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    n = 96
    world = S.team_world()
    even = S.team_split_strided(world, 0, 2, (W + 1) // 2)
    odd = S.team_split_strided(world, 1, 2, W // 2) if W >= 2 else None
    assert S.team_my_pe(world) == me and S.team_n_pes(even) == (W + 1) // 2
    assert S.team_my_pe(even) == (me // 2 if me % 2 == 0 else -1)
    if odd is not None:
        assert S.team_translate_pe(world, 1, odd) == 0 and S.team_translate_pe(world, 0, odd) == -1
        assert S.team_translate_pe(odd, 0, world) == 1
    sync = S.Sync()
    fc = U.nvshmem_create_tensor((W * n,), torch.float32)
    bc = U.nvshmem_create_tensor((n,), torch.float32)
    ring = U.nvshmem_create_tensor((2 * n,), torch.float32)
    sig = U.nvshmem_create_tensor((4,), torch.int32)
    for t in (fc, bc, ring, sig):
        t.zero_()
    U.barrier_all_on_stream()
    for phase in range(1, 4):
        src = torch.arange(n, dtype=torch.float32, device=dev) + 100.0 * me + phase
        S.fcollect(world, sync, fc, src)
        want = torch.cat([torch.arange(n, dtype=torch.float32) + 100.0 * r + phase for r in range(W)])
        _assert_close(fc, want, 0, 0, f"shmem fcollect phase {phase}")
        root = phase % even.size
        S.broadcast(even, sync, bc, src, root)
        if S.team_my_pe(even) >= 0:
            _assert_close(bc, torch.arange(n, dtype=torch.float32) + 100.0 * even.pe(root) + phase, 0, 0, f"shmem broadcast phase {phase}")
        nxt, prev = (me + 1) % W, (me - 1 + W) % W
        S.putmem_signal(ring, src, sig[0:1], phase, S.SIGNAL_SET, nxt)
        S.signal_wait_until(sig[0:1], S.CMP_GE, phase)
        _assert_close(ring[:n], torch.arange(n, dtype=torch.float32) + 100.0 * prev + phase, 0, 0, f"shmem ring phase {phase}")
        S.signal_op(sig[1:2], 1, S.SIGNAL_ADD, 0)                 # everyone adds 1 on PE 0
        S.barrier_all(sync)
        if me == 0:
            if sig.is_cuda:
                torch.cuda.synchronize()
            assert int(sig[1].item()) == W * phase, (sig, phase)
            if not sig.is_cuda:
                assert S.signal_wait_until(sig[1:2], S.CMP_GT, W * phase - 1) == W * phase
        got = torch.empty(n, dtype=torch.float32, device=dev)
        S.getmem(got, ring, nxt)                                  # what I wrote into my successor
        _assert_close(got, src, 0, 0, f"shmem getmem phase {phase}")
        S.barrier_all(sync)
    if fc.is_cuda:
        # ---- the device header, one kernel ----
        from triton_dist import jit
        lib = jit.compile_cuda(SHMEM_TEST_SRC, name="shmem_selftest")
        k = jit.JitKernel(lib, "launch_shmem_selftest")
        slots = U.nvshmem_create_tensor((2 * W,), torch.int32)
        sig64 = U.nvshmem_create_tensor((2,), torch.int64)
        misc = U.nvshmem_create_tensor((4,), torch.int32)
        epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        for t in (slots, sig64, misc, fc, bc, ring):
            t.zero_()
        U.barrier_all_on_stream()
        ctx = jit.symm_ctx()
        for phase in range(1, 4):
            src = torch.arange(n, dtype=torch.float32, device=dev) * 0.5 + 7.0 * me + phase
            k(ctx, slots, epoch, fc, bc, ring, sig64, misc, src, n, phase)
            torch.cuda.synchronize()
            want = torch.cat([torch.arange(n, dtype=torch.float32) * 0.5 + 7.0 * r + phase for r in range(W)])
            _assert_close(fc, want, 0, 0, f"shmem.cuh fcollect phase {phase}")
            root = phase % even.size
            if me % 2 == 0:
                _assert_close(bc, torch.arange(n, dtype=torch.float32) * 0.5 + 7.0 * even.pe(root) + phase, 0, 0, f"shmem.cuh broadcast {phase}")
            prev = (me - 1 + W) % W
            pv = torch.arange(n, dtype=torch.float32) * 0.5 + 7.0 * prev + phase
            _assert_close(ring[:n], pv, 0, 0, f"shmem.cuh putmem_signal phase {phase}")
            _assert_close(ring[n:n + 3], pv[:3], 0, 0, f"shmem.cuh putmem_warp phase {phase}")
            m = misc.cpu().tolist()
            assert m[0] == (me // 2 if me % 2 == 0 else -1) and m[1] == m[0] and m[2] == prev * 10 + phase, (me, m)
            assert int(epoch.item()) == 3 * phase and int(sig64[0].item()) == phase
            U.barrier_all_on_stream()
        for t in (misc, sig64, slots):
            U.nvshmem_free_tensor_sync(t)
    sync.finalize()
    for t in (sig, ring, bc, fc):
        U.nvshmem_free_tensor_sync(t)


def case_lk_shmem():
    """The OpenSHMEM-style device API from a DSL kernel (triton_dist.lk.shmem -> csrc/td/shmem.cuh): the same self-test as the CUDA one in
    case_shmem, written in Python.  GPU backend: generated CUDA; emulation: the interpreter runs transfers, signals and the arrival-flag
    barriers on the shared-memory heap across processes."""
    from triton_dist import lk
    from triton_dist.lk.kernels import simt as K
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    n = 48
    fc = U.nvshmem_create_tensor((W * n,), torch.float32)
    bc = U.nvshmem_create_tensor((n,), torch.float32)
    ring = U.nvshmem_create_tensor((2 * n,), torch.float32)
    slots = U.nvshmem_create_tensor((2 * W,), torch.int32)
    sig64 = U.nvshmem_create_tensor((2,), torch.int64)
    misc = U.nvshmem_create_tensor((4,), torch.int32)
    epoch = torch.zeros(1, dtype=torch.int32, device=dev)
    for t in (slots, sig64, misc, fc, bc, ring):
        t.zero_()
    U.barrier_all_on_stream()
    ctx = lk.symm_ctx()
    n_even = (W + 1) // 2
    for phase in range(1, 4):
        src = torch.arange(n, dtype=torch.float32, device=dev) * 0.5 + 7.0 * me + phase
        args = (ctx, slots, epoch, fc, bc, ring, sig64, misc, src, n, phase)
        if dev.type == "cuda":
            K.shmem_selftest[1](*args)
            torch.cuda.synchronize()
        else:
            K.shmem_selftest.interpret(1, *args)
        want = torch.cat([torch.arange(n, dtype=torch.float32) * 0.5 + 7.0 * r + phase for r in range(W)])
        _assert_close(fc, want, 0, 0, f"lk.shmem fcollect phase {phase}")
        root_pe = 2 * (phase % n_even)
        if me % 2 == 0:
            _assert_close(bc, torch.arange(n, dtype=torch.float32) * 0.5 + 7.0 * root_pe + phase, 0, 0, f"lk.shmem broadcast {phase}")
        prev = (me - 1 + W) % W
        pv = torch.arange(n, dtype=torch.float32) * 0.5 + 7.0 * prev + phase
        _assert_close(ring[:n], pv, 0, 0, f"lk.shmem putmem_signal phase {phase}")
        _assert_close(ring[n:n + 3], pv[:3], 0, 0, f"lk.shmem putmem_warp phase {phase}")
        m = misc.cpu().tolist()
        assert m[0] == (me // 2 if me % 2 == 0 else -1) and m[1] == m[0] and m[2] == prev * 10 + phase, (me, m)
        assert int(epoch.item()) == 3 * phase and int(sig64[0].item()) == phase, (epoch, sig64)
        U.barrier_all_on_stream()
    # ---- the remaining scopes / spellings ----
    n2 = 16
    buf = U.nvshmem_create_tensor((6 * n2 + 2 * W * n2,), torch.int32)
    got = torch.zeros(4 * n2, dtype=torch.int32, device=dev)
    for t in (buf, sig64, misc, slots):
        t.zero_()
    epoch.zero_()
    U.barrier_all_on_stream()
    for phase in range(1, 3):
        src2 = (torch.arange(n2, dtype=torch.int32) + 1000 * me + 10 * phase).to(dev)
        args = (ctx, slots, epoch, buf, got, sig64, misc, src2, n2, phase)
        if dev.type == "cuda":
            K.shmem_selftest_scopes[1](*args)
            torch.cuda.synchronize()
        else:
            K.shmem_selftest_scopes.interpret(1, *args)
        row = lambda r: torch.arange(n2, dtype=torch.int32) + 1000 * r + 10 * phase
        prv, pprv, nxt = (me - 1 + W) % W, (me - 2 + 2 * W) % W, (me + 1) % W
        for reg in range(3):
            assert torch.equal(buf[reg * n2:(reg + 1) * n2].cpu(), row(prv)), ("put", reg, phase)
            assert torch.equal(got[reg * n2:(reg + 1) * n2].cpu(), row(pprv)), ("get", reg, phase)
        assert torch.equal(got[3 * n2:4 * n2].cpu(), row(me)), "remote_ptr loads what I stored on my successor"
        assert torch.equal(buf[3 * n2:4 * n2].cpu(), row(phase % W)), "broadcastmem_block"
        assert torch.equal(buf[6 * n2:6 * n2 + W * n2].cpu(), torch.cat([row(r) for r in range(W)])), "fcollect_warp"
        m = misc.cpu().tolist()
        assert m[:3] == [W - 1, W, me] and int(sig64[1].item()) == phase and (me != 0 or int(sig64[0].item()) == W * phase), (m, sig64)
        U.barrier_all_on_stream()
    U.nvshmem_free_tensor_sync(buf)
    for t in (misc, sig64, slots, ring, bc, fc):
        U.nvshmem_free_tensor_sync(t)


def case_lk_ep():
    """Expert-parallel dispatch / combine as two DSL kernels (warp per (token, k), local slot counters, count + flag publication after a
    grid barrier; pull-combine with fp32 accumulation): dispatch -> every expert scales its rows by (expert id + 1) -> combine, against
    the dense formula.  Three calls reuse the phase-numbered flags; one call overflows the per-source capacity on purpose."""
    from triton_dist.lk.kernels.ep_a2a import LkEpAllToAll
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    gpu = dev.type == "cuda"
    T, H, topk, epr = (96, 256, 4, 2) if gpu else (10, 32, 2, 2)
    E = W * epr
    ep = LkEpAllToAll(T, H, topk, E)
    for it in range(3):
        g = torch.Generator().manual_seed(100 * it + me)
        x = (torch.randn(T, H, generator=g) * 0.5).to(torch.bfloat16).to(dev)
        ids = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32)
        ids[it % T, 0] = -1                                        # one dropped slot
        w = torch.rand(T, topk, generator=g).to(dev)
        recv_x, meta, cnt = ep.dispatch(x, ids.to(dev))
        if gpu:
            torch.cuda.synchronize()
        cnt_h = cnt.cpu().tolist()
        # what every source sends me: pairs whose expert lives here
        all_ids = [torch.zeros(T, topk, dtype=torch.int32, device=dev) for _ in range(W)]
        torch.distributed.all_gather(all_ids, ids.to(dev))
        for src in range(W):
            want = int(((all_ids[src].cpu() >= me * epr) & (all_ids[src].cpu() < (me + 1) * epr)).sum())
            assert cnt_h[src] == want, (it, me, src, cnt_h, want)
        # the experts: row r of source s belongs to local expert meta[s, r, 1] -> scale by (global expert id + 1)
        ep.y_buf.zero_()
        for src in range(W):
            n = cnt_h[src]
            if n:
                scale = (meta[src, :n, 1] + me * epr + 1).to(torch.float32)[:, None]
                ep.y_buf[src, :n] = (recv_x[src, :n].float() * scale).to(torch.bfloat16)
        out = ep.combine(w)
        valid = (ids >= 0).to(dev)
        coef = (w * (ids.to(dev).clamp(min=0) + 1).float() * valid).sum(-1, keepdim=True)
        # bf16 rounding of every expert output, then an fp32 sum: compare against the same two-step formula
        ref = sum(((x.float() * (ids[:, k].to(dev).clamp(min=0) + 1).float()[:, None]).to(torch.bfloat16).float() * (w[:, k] * valid[:, k])[:, None])
                  for k in range(topk))
        _assert_close(out, ref.cpu(), 2e-2, 2e-2, f"lk ep combine iteration {it}")
        assert coef.shape == (T, 1)
    ep.finalize()
    # capacity overflow: with room for one row per (source, destination) the extra pairs are dropped, not written out of bounds
    ep2 = LkEpAllToAll(4, 32 if not gpu else 256, 1, E, cap=1)
    x = torch.ones(4, ep2.H, dtype=torch.bfloat16, device=dev)
    ids = torch.zeros(4, 1, dtype=torch.int32, device=dev)         # everyone routes all 4 tokens to expert 0 (rank 0)
    _, _, cnt = ep2.dispatch(x, ids)
    if gpu:
        torch.cuda.synchronize()
    if me == 0:
        assert cnt.cpu().tolist() == [1] * W, cnt
    assert sorted(ep2.send_slot[:4].cpu().tolist()) == [-1, -1, -1, 0]
    U.barrier_all_on_stream()
    ep2.finalize()


def case_lk_rs_ring():
    """Ring reduce-scatter written in the DSL (W - 1 hops, per-CTA flags carrying the call number, parity double-buffered slots) against
    torch.distributed.reduce_scatter; several calls back to back and two chunk lengths reuse the same buffers."""
    from triton_dist.lk.kernels.reduce_scatter_ring import LkRingReduceScatter
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    max_chunk = 1 << 16 if big else 96
    rs = LkRingReduceScatter(max_chunk)
    for it, chunk in enumerate((max_chunk, max_chunk // 2 + 1, max_chunk, 7)):
        g = torch.Generator().manual_seed(17 * it + me)
        x = torch.randn(W, chunk, generator=g).to(dev)
        out = rs(x)
        ref = torch.empty(chunk, device=dev)
        dist.reduce_scatter(ref, [x[r].contiguous() for r in range(W)], group=U.get_triton_dist_world())
        _assert_close(out, ref, 1e-5, 1e-5, f"lk ring reduce-scatter call {it} chunk {chunk}")
    U.barrier_all_on_stream()
    rs.finalize()


def case_lk_ar_tree():
    """Double-binary-tree all-reduce written in the DSL (up pass + down pass over two complementary trees, per-CTA call-numbered flags)
    against torch.distributed.all_reduce; back-to-back calls and several message lengths on the same buffers (power-of-two worlds)."""
    from triton_dist.lk.kernels.allreduce_tree import LkDoubleTreeAllReduce
    W, me = U.world_size(), U.rank()
    if W & (W - 1):
        return
    dev = U.current_device()
    big = dev.type == "cuda"
    n_max = 1 << 17 if big else 160
    ar = LkDoubleTreeAllReduce(n_max)
    for it, n in enumerate((n_max, n_max // 2 + 2, n_max, 6)):
        g = torch.Generator().manual_seed(31 * it + me)
        x = torch.randn(n, generator=g).to(dev)
        out = ar(x)
        ref = x.clone()
        dist.all_reduce(ref, group=U.get_triton_dist_world())
        _assert_close(out, ref, 1e-5, 1e-5, f"lk double-tree all-reduce call {it} n {n}")
    U.barrier_all_on_stream()
    ar.finalize()


def case_lk_ar_push():
    """One-shot and two-shot push all-reduce written in the DSL against torch.distributed.all_reduce: back-to-back calls, several lengths."""
    from triton_dist.lk.kernels.allreduce_push import LkPushAllReduce
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    n_max = (1 << 16) * W if big else 40 * W
    for method in ("one_shot", "two_shot"):
        ar = LkPushAllReduce(n_max, method)
        for it, n in enumerate((n_max, 3 * W, n_max, W)):
            g = torch.Generator().manual_seed(13 * it + me)
            x = torch.randn(n, generator=g).to(dev)
            out = ar(x).clone()
            ref = x.clone()
            dist.all_reduce(ref, group=U.get_triton_dist_world())
            _assert_close(out, ref, 1e-5, 1e-5, f"lk {method} push all-reduce call {it} n {n}")
        U.barrier_all_on_stream()
        ar.finalize()


def case_lk_ag_ll():
    """Low-latency all-gather with the flag inside the data (8-byte atoms, no barrier, no separate flags) written in the DSL, against
    torch.distributed.all_gather; payloads of several dtypes, back-to-back calls (parity double-buffering)."""
    from triton_dist.lk.kernels.allgather_ll import LkLLAllGather
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    ag = LkLLAllGather(1 << 16 if big else 256)
    for it, (n, dt) in enumerate(((64 if not big else 16384, torch.float32), (10, torch.bfloat16), (33, torch.int32), (64 if not big else 16384, torch.float32))):
        g = torch.Generator().manual_seed(7 * it + me)
        x = (torch.randn(n, generator=g) * 100).to(dt).to(dev)
        out = ag(x)
        ref = [torch.empty_like(x) for _ in range(W)]
        dist.all_gather(ref, x, group=U.get_triton_dist_world())
        assert torch.equal(out.cpu(), torch.stack(ref).cpu()), (it, n, dt)
    U.barrier_all_on_stream()
    ag.finalize()


def case_lk_ar_nvls():
    """NVLS all-reduce written in the DSL (one-shot: every rank reduces the whole message with multimem.ld_reduce; two-shot: every rank reduces
    its share and multimem.st-broadcasts it) against torch.distributed.all_reduce, bf16 and fp32.  On the emulation backend the multicast
    alias is modelled: a multimem load sums the word over every rank's copy, a multimem store writes every copy."""
    from triton_dist.lk.kernels.allreduce_nvls import LkNvlsAllReduce
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    if big and not U.is_nvshmem_multimem_supported():
        return
    nmax = (1 << 18) if big else 16 * 24
    for method in ("one_shot", "two_shot"):
        ar = LkNvlsAllReduce(nmax, method, grid=2 if not big else 4)
        for it, (n, dt) in enumerate(((nmax // 4, torch.float32), (nmax // 2, torch.bfloat16), (8, torch.float32), (nmax // 2, torch.bfloat16))):
            g = torch.Generator().manual_seed(11 * it + me)
            x = (torch.randn(n, generator=g) * 0.5).to(dt).to(dev)
            out = ar(x)
            ref = x.float().clone()
            dist.all_reduce(ref, group=U.get_triton_dist_world())
            _assert_close(out.float(), ref, 5e-2 if dt == torch.bfloat16 else 1e-5, 2e-2 if dt == torch.bfloat16 else 1e-5, f"lk nvls {method} call {it}")
        U.barrier_all_on_stream()
        ar.finalize()


def case_lk_gemm_ar():
    """GEMM + AllReduce as one DSL kernel: tcgen05 tiles stage their output and flag every rank, consumer CTAs reduce finished tiles through
    the multicast alias (multimem.ld_reduce).  Emulation: pipeline model + multicast model across processes; ragged M, three calls."""
    from triton_dist.lk.kernels.gemm_ar import LkGemmArContext, run_gemm_ar
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    gpu = dev.type == "cuda"
    if gpu and not U.is_nvshmem_multimem_supported():
        return
    max_M, K, N = (512, 512, 768) if gpu else (160, 128, 136)
    ctx = LkGemmArContext(max_M, N, BN=128, STAGES=2 if not gpu else 4, N_COMM=2 if not gpu else 8)
    for it, M in enumerate((max_M, max_M - 27, 16)):
        g = torch.Generator().manual_seed(50 * it + me)
        a = (torch.randn(M, K, generator=g) * 0.3).to(torch.bfloat16).to(dev)
        b = (torch.randn(N, K, generator=g) * 0.3).to(torch.bfloat16).to(dev)
        out = run_gemm_ar(ctx, a, b)
        if gpu:
            torch.cuda.synchronize()
        ref = (a.float() @ b.float().t()).to(torch.bfloat16).float()          # every rank's tile is rounded to bf16 before the switch adds
        dist.all_reduce(ref, group=U.get_triton_dist_world())
        _assert_close(out.float(), ref, 6e-2, 3e-2, f"lk gemm_ar call {it} M {M}")
    U.barrier_all_on_stream()
    ctx.finalize()


def case_allreduce_dsl():
    """``TD_ALLREDUCE_DSL=1``: ``ops.comm.all_reduce`` served by the NVLS kernels written in the DSL (A/B switch against the CUDA kernels);
    one- and two-shot methods, bf16 and fp32, a message longer than the workspace (chunked)."""
    from triton_dist.ops import comm
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    if dev.type == "cuda" and not U.is_nvshmem_multimem_supported():
        return
    os.environ["TD_ALLREDUCE_DSL"] = "1"
    try:
        ctx = comm.create_allreduce_ctx(1024, me, W, W)
        for it, (n, dt, method) in enumerate(((128, torch.float32, comm.AllReduceMethod.OneShot_Multimem), (256, torch.bfloat16, comm.AllReduceMethod.TwoShot_Multimem),
                                              (1000, torch.float32, comm.AllReduceMethod.Unknown))):
            g = torch.Generator().manual_seed(5 * it + me)
            x = (torch.randn(n, generator=g) * 0.5).to(dt).to(dev)
            out = comm.all_reduce(x, method, ctx)
            ref = x.float().clone()
            dist.all_reduce(ref, group=U.get_triton_dist_world())
            _assert_close(out.float(), ref, 5e-2 if dt == torch.bfloat16 else 1e-5, 2e-2 if dt == torch.bfloat16 else 1e-5, f"all_reduce via DSL call {it}")
        assert set(ctx._dsl) == {"one_shot", "two_shot"}
        ctx.finalize()
    finally:
        os.environ.pop("TD_ALLREDUCE_DSL", None)


def case_lk_sp_decode():
    """KV-sharded decode entirely on DSL kernels: local split-KV decode, flag-in-data all-gather of the (lse, o) partials, log-sum-exp merge
    across ranks -- against attention over the concatenated cache (one sequence has no keys on the last rank)."""
    from triton_dist.lk.kernels.flash_decode import LkSpDecode
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    B, Hq, Hkv, Lmax = (4, 16, 4, 256) if big else (2, 2, 1, 12)
    sp = LkSpDecode(B, Hq, Hkv, n_splits=2)
    g = torch.Generator().manual_seed(77)                                     # same data on every rank, sliced per rank below
    q = (torch.randn(B, Hq, 128, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    k_all = (torch.randn(B, W * Lmax, Hkv, 128, generator=g) * 0.5).to(torch.bfloat16)
    v_all = (torch.randn(B, W * Lmax, Hkv, 128, generator=g) * 0.5).to(torch.bfloat16)
    total = torch.tensor([W * Lmax - 3] + [max(1, (W - 1) * Lmax - 2)] * (B - 1))     # sequences 1.. end before the last rank's shard
    lens_local = (total - me * Lmax).clamp(0, Lmax).to(torch.int32).to(dev)
    for it in range(2):
        out = sp(q, k_all[:, me * Lmax:(me + 1) * Lmax].contiguous().to(dev), v_all[:, me * Lmax:(me + 1) * Lmax].contiguous().to(dev), lens_local)
        G = Hq // Hkv
        ref = torch.zeros(B, Hq, 128)
        for b in range(B):
            n = int(total[b])
            for h in range(Hq):
                s_ = (q[b, h].float().cpu() @ k_all[b, :n, h // G].float().t()) * 128 ** -0.5
                ref[b, h] = torch.softmax(s_, -1) @ v_all[b, :n, h // G].float()
        _assert_close(out.float(), ref, 2e-2, 2e-2, f"lk sp decode call {it}")
    U.barrier_all_on_stream()
    sp.finalize()


def case_lk_a2a():
    """Low-latency variable-size all-to-all written with the DSL's OpenSHMEM-style API (put + put-with-signal per destination CTA, wait on the
    source's signal) against torch.distributed.all_to_all_single with uneven splits; three calls on double-buffered slots."""
    from triton_dist.lk.kernels.all_to_all import LkAllToAll
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    max_rows, H = (256, 512) if big else (6, 16)
    a2a = LkAllToAll(max_rows, H)
    for it in range(3):
        g = torch.Generator().manual_seed(9 * it + me)
        splits = torch.randint(0, max_rows + 1, (W,), generator=g).to(torch.int32)
        send = (torch.randn(int(splits.sum()), H, generator=g)).to(torch.bfloat16).to(dev)
        recv, cnt = a2a(send, splits.to(dev))
        if big:
            torch.cuda.synchronize()
        all_splits = [torch.zeros(W, dtype=torch.int32, device=dev) for _ in range(W)]
        dist.all_gather(all_splits, splits.to(dev), group=U.get_triton_dist_world())
        want_cnt = [int(all_splits[s][me]) for s in range(W)]
        assert cnt.cpu().tolist() == want_cnt, (it, cnt, want_cnt)
        out = torch.empty(sum(want_cnt), H, dtype=torch.float32, device=dev)          # bf16 values are exact in fp32 (gloo has no bf16 / int16)
        dist.all_to_all_single(out, send.float(), want_cnt, splits.tolist(), group=U.get_triton_dist_world())
        off = 0
        for s_ in range(W):
            assert torch.equal(recv[s_, :want_cnt[s_]].float().cpu(), out[off:off + want_cnt[s_]].cpu()), (it, s_)
            off += want_cnt[s_]
        U.barrier_all_on_stream()
    a2a.finalize()


def case_lk_nvls_collectives():
    """Reduce-scatter (multimem.ld_reduce of the own chunk) and all-gather (one multimem.st stream per rank) through the multicast alias,
    written in the DSL, against torch.distributed; bf16 and fp32, alternating calls on one staging buffer."""
    from triton_dist.lk.kernels.collectives_nvls import LkNvlsCollectives
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    if big and not U.is_nvshmem_multimem_supported():
        return
    unit = (1 << 14) if big else 16                        # elements per chunk / shard
    col = LkNvlsCollectives(unit * W * 4, grid=2 if not big else 4)
    for it, dt in enumerate((torch.float32, torch.bfloat16, torch.float32)):
        g = torch.Generator().manual_seed(3 * it + me)
        x = (torch.randn(W * unit, generator=g) * 0.5).to(dt).to(dev)
        rs = col.reduce_scatter(x)
        full = x.float().clone()
        dist.all_reduce(full, group=U.get_triton_dist_world())
        _assert_close(rs.float(), full[me * unit:(me + 1) * unit], 5e-2 if dt == torch.bfloat16 else 1e-5, 2e-2 if dt == torch.bfloat16 else 1e-5, f"lk nvls reduce_scatter {it}")
        shard = x[:unit].contiguous()
        ag = col.all_gather(shard)
        ref = [torch.empty(unit, dtype=torch.float32, device=dev) for _ in range(W)]
        dist.all_gather(ref, shard.float(), group=U.get_triton_dist_world())
        assert torch.equal(ag.float().cpu(), torch.stack(ref).cpu()), it
    U.barrier_all_on_stream()
    col.finalize()


def case_lk_ag_gemm():
    """AllGather + GEMM as ONE kernel written in the Python DSL (comm CTAs push shards + release-add flags, tcgen05 tiles acquire the
    flags of the rows they need).  GPU: the generated CUDA; emulation: the interpreter with the functional pipeline model, across ranks."""
    from triton_dist.lk.kernels.ag_gemm import LkAgGemmContext, run_ag_gemm
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    gpu = dev.type == "cuda"
    Ms, K, N = (256, 512, 768) if gpu else (128, 128, 256)
    ctx = LkAgGemmContext(Ms, K, BN=256, STAGES=2 if not gpu else 4, N_COMM=2 if not gpu else 4)
    g = torch.Generator().manual_seed(21)
    b = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    for it in range(3 if gpu else 2):
        ga = torch.Generator().manual_seed(1000 * it + me)
        a = (torch.randn(Ms, K, generator=ga) * 0.5).to(torch.bfloat16).to(dev)
        out = run_ag_gemm(ctx, a, b)
        full = torch.empty(W * Ms, K, dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(full, a, group=U.get_triton_dist_world())
        _assert_close(out, full.float() @ b.float().t(), 0.25, 2e-2, f"lk ag_gemm call {it}")
    ctx.finalize()


def case_lk_gemm_rs():
    """GEMM + ReduceScatter as ONE DSL kernel: tcgen05 tiles reduce straight into the owner's buffer (16-byte bf16x2 reductions over
    symm_at addresses), per-tile release-add of the owner's counter, collector CTAs acquire it.  GPU: generated CUDA; emulation: the
    interpreter across ranks (pipeline model + CAS-emulated reductions)."""
    from triton_dist.lk.kernels.gemm_rs import LkGemmRsContext, run_gemm_rs
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    gpu = dev.type == "cuda"
    Ms, K, N = (256, 512, 768) if gpu else (128, 128, 256)
    ctx = LkGemmRsContext(Ms, N, BN=256, STAGES=2 if not gpu else 4, N_COLLECT=2 if not gpu else 4)
    for it in range(3 if gpu else 2):
        g = torch.Generator().manual_seed(500 * it + me)
        a = (torch.randn(W * Ms, K, generator=g) * 0.25).to(torch.bfloat16).to(dev)
        b = (torch.randn(N, K, generator=g) * 0.25).to(torch.bfloat16).to(dev)
        out = run_gemm_rs(ctx, a, b)
        part = a.float() @ b.float().t()                      # my K shard's contribution to all rows
        dist.all_reduce(part, group=U.get_triton_dist_world())
        _assert_close(out, part[me * Ms:(me + 1) * Ms], 0.3, 3e-2, f"lk gemm_rs call {it}")
    ctx.finalize()


def case_ep_fn_api():
    """The process-wide EP op API of the autograd functions (reference: function/nvidia/common.py): init_triton_dist_ep_op ->
    init_triton_dist_ep_ctx -> fused_ep_moe (forward vs the golden; forward + backward on the CUDA backend) -> deinit; split_mbs gives
    two ops / streams."""
    from triton_dist.function import nvidia as F
    from triton_dist.ops import ep_mega as EM
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    bf = torch.bfloat16 if big else torch.float32
    T, H, I, epr, topk = (256, 256, 128, 2, 2) if big else (12, 16, 8, 2, 2)
    E = epr * W
    grp = U.get_triton_dist_world()
    assert not F.triton_dist_ep_op_initialized("mega")
    op = F.init_triton_dist_ep_op(grp, T, H, topk, me, E, W, dtype=bf, num_sm=16, capacity=3.0, ep_implementation="mega")
    assert F.triton_dist_ep_op_initialized("mega") and F.get_triton_dist_ep_op(0) is op and F.get_ep_capacity("mega") == 3.0
    ectx = F.init_triton_dist_ep_ctx(grp, topk, E, "mega")
    assert ectx.ep_op is op and ectx.num_experts_per_rank == epr and ectx.max_num_tiles >= epr
    cfg = F.get_moe_optim_config(use_mega=True, is_forward=True)
    assert isinstance(cfg, F.MoEOptimConfig) and cfg.num_dispatch_sms > 0
    g = torch.Generator(device="cpu").manual_seed(3)
    w_gu_all = (torch.randn(E, 2 * I, H, generator=g) * 0.05).to(bf).to(dev)
    w_dn_all = (torch.randn(E, H, I, generator=g) * 0.05).to(bf).to(dev)
    w_gu, w_dn = w_gu_all[me * epr:(me + 1) * epr].contiguous(), w_dn_all[me * epr:(me + 1) * epr].contiguous()
    for it in range(2):
        x = (torch.randn(T, H, device=dev) * 0.5).to(bf)
        ids = torch.randn(T, E, device=dev).topk(topk, dim=1).indices.to(torch.int32)
        wts = torch.softmax(torch.randn(T, topk, device=dev), -1)
        if big:
            xg, wg = x.clone().requires_grad_(True), w_gu.clone().requires_grad_(True)
            out = F.fused_ep_moe(xg, ids, wts, wg, w_dn, ectx)
            out.float().sum().backward()
            assert xg.grad is not None and wg.grad is not None and torch.isfinite(xg.grad.float()).all()
        else:
            out = EM.mega_ep_moe(ectx.ep_op, x, ids, wts, w_gu, w_dn)          # the emulation backend mirrors the forward protocol
        ref = EM.mega_ep_moe_reference(x, ids, wts, w_gu_all, w_dn_all)
        _assert_close(out.detach(), ref, 0.05 if big else 1e-4, 5e-2 if big else 1e-4, f"fused_ep_moe it{it}")
    U.barrier_all_host()
    F.deinit_triton_dist_ep_op("mega")
    assert not F.triton_dist_ep_op_initialized("mega")
    F.init_triton_dist_ep_op(grp, T, H, topk, me, E, W, dtype=bf, capacity=3.0, ep_implementation="split_mbs")
    assert F.triton_dist_ep_op_initialized("split_mbs") and F.get_triton_dist_ep_op(1) is not F.get_triton_dist_ep_op(2)
    c1, c2 = F.init_triton_dist_ep_ctx(grp, topk, E, "split_mbs", 0), F.init_triton_dist_ep_ctx(grp, topk, E, "split_mbs", 1)
    assert c1.ep_op is F.get_triton_dist_ep_op(1) and c2.ep_op is F.get_triton_dist_ep_op(2)
    U.barrier_all_host()
    F.deinit_triton_dist_ep_op("split_mbs")
    F.set_triton_dist_moe_profile_enabled(True, "/tmp/td_prof")
    assert F.get_triton_dist_moe_profile_enabled()["enabled"] and F.get_triton_dist_profile_output_dir() == "/tmp/td_prof"
    F.set_triton_dist_moe_profile_enabled(False)


def case_mega():
    """Megakernel decode step (task graph + scoreboard + in-kernel all-reduce) vs the layer-by-layer TP model."""
    from triton_dist.models import AutoLLM, KV_Cache, ModelConfig
    from triton_dist.mega_kernel import MegaDenseModel
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    cfg = ModelConfig(model_name="tiny-dense", max_length=64, dtype=dtype, rank=me, world_size=W)
    m = AutoLLM.from_pretrained(cfg, U.get_triton_dist_world())
    B = 2
    mk = lambda: KV_Cache(m.num_layers, B, 64, m.num_key_value_heads, m.head_dim, dtype, W, dev)
    kv, kv2 = mk(), mk()
    g = torch.Generator(device="cpu").manual_seed(3 + me)
    kv.k_cache.copy_((torch.randn(kv.k_cache.shape, generator=g) * 0.5).to(dtype)); kv.v_cache.copy_((torch.randn(kv.v_cache.shape, generator=g) * 0.5).to(dtype))
    kv.kv_offset.fill_(9)
    kv2.k_cache.copy_(kv.k_cache); kv2.v_cache.copy_(kv.v_cache); kv2.kv_offset.copy_(kv.kv_offset)
    mega = MegaDenseModel(m, B, kv2, attn_splits=3)
    for step in range(3):
        ids = torch.randint(0, 1000, (B, 1), generator=torch.Generator().manual_seed(40 + step)).to(dev)
        pos = kv.kv_offset.to(torch.int64)[:, None]
        m.set_fwd("torch")
        ref = m.inference(ids, pos, kv)
        out = mega.mega_forward(ids)
        _assert_close(out, ref, 6e-2 if big else 1e-4, 6e-2 if big else 1e-4, f"megakernel logits step {step}")
        kv.inc_offset(1); kv2.inc_offset(1)
    U.barrier_all_host()
    mega.finalize()


def case_ep_metadata():
    """Routing metadata ops of the throughput-mode EP path (reference ep_a2a_intra_node.py:423): all-gathered per-expert histograms ->
    receive offsets / token counts, and the per-token rows relative to the destination rank's buffer."""
    from triton_dist.ops import ep_metadata as EM
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    epr, T, topk = 2, 11 + U.rank(), 2
    E = W * epr
    g = torch.Generator().manual_seed(5 + me)
    idx = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32).to(dev)
    offs, n_recv, n_in, full = EM.get_ag_splits_and_recv_offset_for_dispatch_intra_node(idx, E)
    hists = [torch.zeros(E + 1, dtype=torch.int32, device=dev) for _ in range(W)]
    torch.distributed.all_gather(hists, EM.expert_histogram(idx, E))
    want = torch.stack(hists)
    assert torch.equal(full.cpu(), want.cpu()), (full, want)
    for r in range(W):
        run = 0
        for e in range(epr):
            for s_ in range(W):
                assert int(offs[r, e, s_]) == run
                run += int(want[s_, r * epr + e])
        assert int(n_recv[r]) == run
    assert int(n_in[me]) == T * topk
    U.barrier_all_host()


def case_allgather_ring():
    """Ring copy-engine all-gather producers (1-D: W-1 hops; 2-D: ring inside a group + across groups) on the reference's buffer / flag
    contract (allgather.py:127-200)."""
    from triton_dist.ops.allgather import create_allgather_buffers
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    bufs, flags = create_allgather_buffers(64 * W, 32, torch.float32)
    U.barrier_all_on_stream()
    from triton_dist.ops.allgather import AllGatherMethod, cp_engine_producer_all_gather
    sig = 4
    for method in (AllGatherMethod.Ring1D_IntraNode, AllGatherMethod.Ring2D_IntraNode):
        for it in range(3):
            local = torch.randn(64, 32, device=dev)
            cp_engine_producer_all_gather(me, W, local, bufs, flags, signal_value=sig, method=method)
            if dev.type == "cuda":
                torch.cuda.synchronize()
            ref = torch.empty(64 * W * 32, device=dev)
            dist.all_gather_into_tensor(ref, local.view(-1), group=U.get_triton_dist_world())
            assert torch.equal(bufs[me].cpu().view(-1), ref.cpu()), (method, it)
            U.barrier_all_host()                               # nobody overwrites a buffer a peer is still checking
            sig += 1


def case_ulysses_pack():
    """Ulysses q, k, v in ONE packed all-to-all (``UlyssesQKVPackAllToAll``): seq-sharded [S/W, H, D] -> head-sharded [S, H/W, D]."""
    W, me = U.world_size(), U.rank()
    dev = U.current_device()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    g = torch.Generator().manual_seed(5)
    S_l, H, D = 8, 2 * W, 128
    full = torch.randn(S_l * W, H, D, generator=g).to(dtype).to(dev)
    from triton_dist.parallel.sp import UlyssesQKVPackAllToAll
    Hkv_u = W
    pk = UlyssesQKVPackAllToAll(S_l, H, Hkv_u, D, dtype, me, W)
    kfull = torch.randn(S_l * W, Hkv_u, D, generator=g).to(dtype).to(dev)
    vfull = torch.randn(S_l * W, Hkv_u, D, generator=g).to(dtype).to(dev)
    sl = slice(me * S_l, (me + 1) * S_l)
    for _ in range(3):
        q2, k2, v2 = pk(full[sl].contiguous(), kfull[sl].contiguous(), vfull[sl].contiguous())
        assert torch.equal(q2.cpu(), full[:, me * (H // W):(me + 1) * (H // W)].cpu())
        assert torch.equal(k2.cpu(), kfull[:, me:me + 1].cpu()) and torch.equal(v2.cpu(), vfull[:, me:me + 1].cpu())
    pk.finalize()


def case_mega_paged():
    """Megakernel decode step through a paged KV cache (block-table task types) with TP-sharded heads, vs the layer-by-layer model on a
    dense cache holding the same tokens."""
    from triton_dist.models import AutoLLM, KV_Cache, ModelConfig
    from triton_dist.mega_kernel import MegaDenseModel
    dev = U.current_device()
    W, me = U.world_size(), U.rank()
    big = dev.type == "cuda"
    dtype = torch.bfloat16 if big else torch.float32
    cfg = ModelConfig(model_name="tiny-dense", max_length=64, dtype=dtype, rank=me, world_size=W)
    m = AutoLLM.from_pretrained(cfg, U.get_triton_dist_world())
    m.set_fwd("torch")
    B = 2
    kv = KV_Cache(m.num_layers, B, 64, m.num_key_value_heads, m.head_dim, dtype, W, dev)
    g = torch.Generator(device="cpu").manual_seed(3 + me)
    kv.k_cache.copy_((torch.randn(kv.k_cache.shape, generator=g) * 0.5).to(dtype)); kv.v_cache.copy_((torch.randn(kv.v_cache.shape, generator=g) * 0.5).to(dtype))
    kv.kv_offset.fill_(9)
    from triton_dist.models import PagedKVCache
    L = int(kv.kv_offset[0])
    paged = PagedKVCache(PAGE_SIZE=8, num_layers=m.num_layers, batch_size=B, max_length=64, num_kv_heads=kv.kv_heads, head_dim=m.head_dim,
                         dtype=dtype, device=dev, seed=11 + me)
    for li in range(m.num_layers):
        k, v = kv.layer(li)
        paged.append(li, k[:, :L], v[:, :L])
    paged.inc_offset(L)
    mega_p = MegaDenseModel(m, B, paged, attn_splits=2)
    for step in range(2):
        ids = torch.randint(0, 1000, (B, 1), generator=torch.Generator().manual_seed(70 + step)).to(dev)
        ref = m.inference(ids, kv.kv_offset.to(torch.int64)[:, None], kv)
        out = mega_p.mega_forward(ids)
        _assert_close(out, ref, 6e-2 if big else 1e-4, 6e-2 if big else 1e-4, f"megakernel paged logits step {step}")
        kv.inc_offset(1); paged.inc_offset(1)
    U.barrier_all_host()
    mega_p.finalize()


def case_mega_server():
    """The megakernel text-generation service across ranks: rank 0 serves a socket from a thread and broadcasts every request, the other
    ranks follow; per-op prefill, token-by-token prefill and the paged KV cache must generate the same tokens (greedy and seeded)."""
    import threading
    from triton_dist.mega_kernel.server import Client, MegaServer
    dev = U.current_device()
    me = U.rank()
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    seen = []
    for kw in (dict(prefill="per_op"), dict(prefill="stepwise"), dict(page_size=4)):
        srv = MegaServer("tiny-dense", max_length=64, dtype=dtype, port=0, max_prompt=32, **kw)
        if me == 0:
            ready = threading.Event()
            th = threading.Thread(target=srv.serve_forever, kwargs=dict(ready=ready), daemon=True)
            th.start()
            assert ready.wait(60)
            with Client(port=srv.port) as c:
                r = c.ask("hi there", max_new_tokens=5, temperature=0.0)
                r2 = c.request({"prompt_ids": [1, 2, 3], "max_new_tokens": 4, "seed": 7})
                assert r["status"] == r2["status"] == "success" and r["prompt_tokens"] == 8 and len(r["token_ids"]) == 5, (r, r2)
                assert c.request({"cmd": "stats"})["requests"] == 2 and c.request({"nothing": 1})["status"] == "error"
                assert c.request({"cmd": "shutdown"})["status"] == "success"
            th.join(60)
            assert not th.is_alive()
            seen.append((r["token_ids"], r2["token_ids"]))
        else:
            srv.serve_forever()
        U.barrier_all_host()
        srv.finalize()
    if me == 0 and dev.type != "cuda":          # fp32 emulation: the three prefill routes agree token for token
        assert seen[0] == seen[1] == seen[2], seen


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    U.initialize_distributed(seed=1 + int(os.environ.get("RANK", 0)))
    t0 = time.time()
    for name in names:
        CASES[name]()
        U.barrier_all_host()
        U.dist_print(f"CASE {name} OK ({time.time() - t0:.1f}s)", allowed_ranks=[0])
    U.finalize_distributed()
