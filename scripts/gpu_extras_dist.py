"""Multi-GPU perf sweep: fused ag_gemm / gemm_rs vs NCCL+cuBLAS and vs the GEMM-only twin, per config.
torchrun --nproc-per-node N scripts/gpu_extras_dist.py  -- EP low-latency and fused GEMM+AR latencies"""
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
# EP low-latency dispatch / combine at the reference's published config (README.md:98-99: 128 tokens/rank, topk 8, hidden 7168, fp8)
try:
    from triton_dist.ops import ep_a2a as EP
    T, H, topk, E = 128, 7168, 8, 32 * W
    for fp8 in (True, False):
        ectx = EP.create_ep_ll_a2a_ctx(T, H, topk, E, online_quant_fp8=fp8, dtype=bf)
        g = torch.Generator().manual_seed(me)
        x = (torch.randn(T, H, generator=g) * 0.5).to(bf).to(dev)
        idx = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32).to(dev)
        wts = torch.softmax(torch.randn(T, topk, generator=g), -1).to(dev)
        rx, rs, cnt, meta = EP.ep_ll_dispatch(ectx, x, idx)
        y = torch.zeros((E // W, rx.shape[1], H), dtype=bf, device=dev)
        td = graph_us(lambda: EP.ep_ll_dispatch(ectx, x, idx), 10)
        def both():
            r = EP.ep_ll_dispatch(ectx, x, idx)
            EP.ep_ll_combine(ectx, y, idx, wts, r[3])
        tb = graph_us(both, 10)
        emit(dict(op="ep_ll", W=W, tokens=T, hidden=H, topk=topk, experts=E, fp8=fp8, dispatch_us=td, dispatch_combine_us=tb,
                  ref_note="reference: ~76 + ~126 us on 8xH800 (low_latency_a2a_v2.rst:21-31), 137 us on 32xH800"))
        U.barrier_all_host(); ectx.finalize()
except Exception as e:
    emit(dict(op="ep_ll", error=str(e)[:300]))
# single-kernel GEMM+AllReduce (decode shapes) vs cuBLAS + NCCL all-reduce, under a CUDA graph
try:
    from triton_dist.ops.gemm_ar import create_ll_gemm_ar_context, low_latency_gemm_allreduce_op, gemm_allreduce_op
    for (M, N, K) in ((128, 5120, 25600 // W * 1), (16, 4096, 12288 // W)):
        gctx = create_ll_gemm_ar_context(None, me, W, W, max_M=M, N=N, dtype=bf)
        a = (torch.randn(M, K, device=dev) * 0.1).to(bf); wt = (torch.randn(N, K, device=dev) * 0.1).to(bf)
        o = torch.empty(M, N, device=dev, dtype=bf)
        def nccl_ar():
            torch.matmul(a, wt.t(), out=o); dist.all_reduce(o, group=grp)
        t0 = graph_us(nccl_ar, 10)
        t1 = graph_us(lambda: low_latency_gemm_allreduce_op(gctx, a, wt, out=o), 10)
        t2 = graph_us(lambda: gemm_allreduce_op(gctx, a, wt, out=o), 10)
        emit(dict(op="gemm_ar", M=M, N=N, K=K, W=W, fused_one_kernel_us=t1, two_kernel_us=t2, cublas_nccl_us=t0))
        U.barrier_all_host(); gctx.finalize()
except Exception as e:
    emit(dict(op="gemm_ar", error=str(e)[:300]))
# intra-kernel profiles of the two headline kernels (one call each; Perfetto traces + summaries for offline analysis)
try:
    from triton_dist.tools.profiler import ProfilerBuffer, export_to_perfetto_trace, summarize
    os.makedirs("gpurun_out", exist_ok=True)
    M, N, K = 4096, 12288, 49152
    rctx = create_gemm_rs_context(M, N, output_dtype=bf)
    A = (torch.randn(M, K // W, device=dev) * 0.1).to(bf); Bw = (torch.randn(N, K // W, device=dev) * 0.1).to(bf)
    for _ in range(3): gemm_rs(A, Bw.t(), rctx)
    pb = ProfilerBuffer(); torch.cuda.synchronize(); dist.barrier(group=grp)
    gemm_rs(A, Bw.t(), rctx, profiler=pb); torch.cuda.synchronize()
    if me == 0:
        emit(dict(op="profile_gemm_rs", summary=summarize(pb)))
        export_to_perfetto_trace(pb, f"gpurun_out/gemm_rs_trace_n{W}.json.gz", rank=me)
    U.barrier_all_host(); rctx.finalize(); del A, Bw
    M, N, K = 4096, 4096, 4096
    actx = create_ag_gemm_context(M, N // W, K, bf)
    A = torch.randn(M // W, K, device=dev, dtype=bf); Bw = torch.randn(N // W, K, device=dev, dtype=bf)
    for _ in range(3): ag_gemm(A, Bw.t(), actx)
    pb = ProfilerBuffer(); torch.cuda.synchronize(); dist.barrier(group=grp)
    ag_gemm(A, Bw.t(), actx, profiler=pb); torch.cuda.synchronize()
    if me == 0:
        emit(dict(op="profile_ag_gemm", summary=summarize(pb)))
        export_to_perfetto_trace(pb, f"gpurun_out/ag_gemm_trace_n{W}.json.gz", rank=me)
    U.barrier_all_host(); actx.finalize()
except Exception as e:
    emit(dict(op="profile", error=str(e)[:300]))
if me == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/extras_dist_n{W}.json", "w"), indent=1)
U.finalize_distributed()
