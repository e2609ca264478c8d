"""Multi-GPU perf sweep: fused ag_gemm / gemm_rs vs NCCL+cuBLAS and vs the GEMM-only twin, per config.
torchrun --nproc-per-node N scripts/gpu_sweep_dist.py [quick]"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.ops.ag_gemm import ag_gemm, create_ag_gemm_context
from triton_dist.ops.gemm_rs import create_gemm_rs_context, gemm_rs
from triton_dist.ops.gemm import GemmConfig, gemm
from triton_dist.ops import comm

U.initialize_distributed(seed=0, heap_bytes=6 << 30)
W, me = U.world_size(), U.rank()
dev = U.current_device(); grp = U.get_triton_dist_world()
bf = torch.bfloat16
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
rows = []

def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(group=grp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
    return t.item()

def emit(d):
    rows.append(d)
    if me == 0: print(json.dumps(d), flush=True)

AG_SHAPES = [(4096, 4096, 4096), (8192, 11008, 4096), (8192, 28672, 8192)] if not quick else [(4096, 4096, 4096)]
RS_SHAPES = [(4096, 12288, 49152), (8192, 4096, 11008), (8192, 8192, 28672)] if not quick else [(4096, 12288, 49152)]

for (M, N, K) in AG_SHAPES:
    Nl = N // W
    A = torch.randn(M // W, K, device=dev, dtype=bf) * 0.05
    B = torch.randn(Nl, K, device=dev, dtype=bf) * 0.05
    out = torch.empty(M, Nl, device=dev, dtype=bf)
    full = torch.empty(M, K, device=dev, dtype=bf)
    ctx = create_ag_gemm_context(M, Nl, K, bf)
    def nccl():
        dist.all_gather_into_tensor(full, A, group=grp); torch.matmul(full, B.t(), out=out)
    t_nccl = timed(nccl)
    t_ag_only = timed(lambda: dist.all_gather_into_tensor(full, A, group=grp))
    t_cublas = timed(lambda: torch.matmul(full, B.t(), out=out))
    dist.all_gather_into_tensor(full, A, group=grp)
    ref = full.float() @ B.float().t()
    flops = 2.0 * M * N * K
    emit(dict(op="ag_gemm", M=M, N=N, K=K, W=W, impl="nccl+cublas", ms=t_nccl, nccl_ag_ms=t_ag_only, cublas_ms=t_cublas, tflops=flops / t_nccl / 1e9))
    for cg in (2, 1):
        if (M // W) % (128 * cg): continue
        for bn in (256,):
            if Nl < bn: continue
            t_twin = timed(lambda: gemm(full, B, out=out, config=GemmConfig(bn=bn, cta_group=cg, group_m=8)))
            for nc in (8, 16, 32):
                cfg = GemmConfig(bn=bn, cta_group=cg, group_m=max(1, (M // W) // (128 * cg)), use_tma_store=True, n_comm_ctas=nc)
                try:
                    c = ag_gemm(A, B.t(), ctx, gemm_config=cfg, out=out)
                    err = (c.float() - ref).abs().max().item()
                    t = timed(lambda: ag_gemm(A, B.t(), ctx, gemm_config=cfg, out=out))
                    emit(dict(op="ag_gemm", M=M, N=N, K=K, W=W, impl="ours", cg=cg, bn=bn, n_comm=nc, ms=t, gemm_only_ms=t_twin,
                              exposed_us=(t - t_twin) * 1e3, tflops=flops / t / 1e9, speedup=t_nccl / t, max_err=err))
                except Exception as e:
                    emit(dict(op="ag_gemm", M=M, N=N, K=K, cg=cg, bn=bn, n_comm=nc, error=str(e)[:200]))
            try:
                cfg = GemmConfig(bn=bn, cta_group=cg, group_m=max(1, (M // W) // (128 * cg)), use_tma_store=True, n_comm_ctas=0)
                c = ag_gemm(A, B.t(), ctx, gemm_config=cfg, out=out, transport="copy_engine")
                err = (c.float() - ref).abs().max().item()
                t = timed(lambda: ag_gemm(A, B.t(), ctx, gemm_config=cfg, out=out, transport="copy_engine"))
                emit(dict(op="ag_gemm", M=M, N=N, K=K, W=W, impl="ours(copy_engine)", cg=cg, bn=bn, n_comm=0, ms=t, gemm_only_ms=t_twin,
                          exposed_us=(t - t_twin) * 1e3, tflops=flops / t / 1e9, speedup=t_nccl / t, max_err=err))
            except Exception as e:
                emit(dict(op="ag_gemm", M=M, N=N, K=K, cg=cg, bn=bn, impl="ours(copy_engine)", error=str(e)[:200]))
    U.barrier_all_host(); ctx.finalize()

for (M, N, K) in RS_SHAPES:
    Kl = K // W
    A = torch.randn(M, Kl, device=dev, dtype=bf) * 0.05
    B = torch.randn(N, Kl, device=dev, dtype=bf) * 0.05
    out = torch.empty(M // W, N, device=dev, dtype=bf)
    full = torch.empty(M, N, device=dev, dtype=bf)
    ctx = create_gemm_rs_context(M, N, output_dtype=bf)
    def nccl():
        torch.matmul(A, B.t(), out=full); dist.reduce_scatter_tensor(out, full, group=grp)
    t_nccl = timed(nccl)
    t_cublas = timed(lambda: torch.matmul(A, B.t(), out=full))
    t_rs_only = timed(lambda: dist.reduce_scatter_tensor(out, full, group=grp))
    nccl(); ref = out.float().clone()
    flops = 2.0 * M * N * K
    emit(dict(op="gemm_rs", M=M, N=N, K=K, W=W, impl="nccl+cublas", ms=t_nccl, cublas_ms=t_cublas, nccl_rs_ms=t_rs_only, tflops=flops / t_nccl / 1e9))
    for cg in (2, 1):
        if (M // W) % (128 * cg): continue
        for bn in (256,):
            cfg = GemmConfig(bn=bn, cta_group=cg, group_m=max(1, (M // W) // (128 * cg)), use_tma_store=False)
            t_twin = timed(lambda: gemm(A, B, out=full, config=GemmConfig(bn=bn, cta_group=cg, group_m=8)))
            try:
                c = gemm_rs(A, B.t(), ctx, gemm_config=cfg, out=out)
                err = (c.float() - ref).abs().max().item()
                t = timed(lambda: gemm_rs(A, B.t(), ctx, gemm_config=cfg, out=out))
                emit(dict(op="gemm_rs", M=M, N=N, K=K, W=W, impl="ours", cg=cg, bn=bn, ms=t, gemm_only_ms=t_twin,
                          exposed_us=(t - t_twin) * 1e3, tflops=flops / t / 1e9, speedup=t_nccl / t, max_err=err, ref_max=ref.abs().max().item()))
            except Exception as e:
                emit(dict(op="gemm_rs", M=M, N=N, K=K, cg=cg, bn=bn, error=str(e)[:200]))
    U.barrier_all_host(); ctx.finalize()

# all-reduce latency / bandwidth table
arctx = comm.create_allreduce_ctx(64 << 20, me, W, W)
for n in ([2048, 65536, 1 << 20, 16 << 20, 64 << 20] if not quick else [65536, 16 << 20]):
    x = torch.randn(n // 2, device=dev, dtype=bf)
    o = torch.empty_like(x)
    t_nccl = timed(lambda: dist.all_reduce(x, group=grp), 20, 5)
    for m in (comm.AllReduceMethod.OneShot, comm.AllReduceMethod.TwoShot, comm.AllReduceMethod.OneShot_Multimem, comm.AllReduceMethod.TwoShot_Multimem):
        if "Multimem" in m.name and not U.is_nvshmem_multimem_supported(): continue
        if n > (1 << 20) and m == comm.AllReduceMethod.OneShot: continue
        t = timed(lambda: comm.all_reduce(x, m, arctx, output=o), 20, 5)
        emit(dict(op="all_reduce", bytes=n, W=W, method=m.name, us=t * 1e3, nccl_us=t_nccl * 1e3, busbw_gbs=2 * (W - 1) / W * n / t / 1e6))
# small-message latency without Python launch overhead: 20 calls captured in one CUDA graph
def graph_us(fn, n=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier(group=grp)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    torch.cuda.synchronize(); dist.barrier(group=grp)
    return timed(g.replay, 10, 3) * 1e3 / n
for n in (2048, 8192, 65536, 1 << 20):
    x = torch.randn(n // 2, device=dev, dtype=bf); o = torch.empty_like(x)
    try:
        t_nccl = graph_us(lambda: dist.all_reduce(x, group=grp))
    except Exception as e:
        t_nccl = float("nan")
    for m in (comm.AllReduceMethod.OneShot, comm.AllReduceMethod.OneShot_Multimem, comm.AllReduceMethod.TwoShot_Multimem):
        if "Multimem" in m.name and not U.is_nvshmem_multimem_supported(): continue
        t = graph_us(lambda: comm.all_reduce(x, m, arctx, output=o))
        emit(dict(op="all_reduce_graph", bytes=n, W=W, method=m.name, us=t, nccl_us=t_nccl))
if me == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/sweep_dist_n{W}.json", "w"), indent=1)
U.finalize_distributed()
