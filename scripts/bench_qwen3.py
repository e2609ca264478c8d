"""Qwen3-8B-shaped TP decode step (bs=1, ctx=512 by default), random-init weights: ms/step and tokens/s per backend.
Mirrors the reference's megakernel doc table (docs/getting-started/megakernel/megakernel.md:31-34):
torch eager / torch + CUDA graph / triton_dist_AR + graph / gemm_ar + graph / MegaKernel.
  torchrun --nproc-per-node W scripts/bench_qwen3.py [--model Qwen/Qwen3-8B] [--bs 1] [--ctx 512] [--layers N]"""
import argparse, json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist.models import AutoLLM, KV_Cache, ModelConfig
from triton_dist.mega_kernel import MegaDenseModel

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="Qwen/Qwen3-8B"); ap.add_argument("--bs", type=int, default=1); ap.add_argument("--ctx", type=int, default=512)
ap.add_argument("--layers", type=int, default=0); ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
U.initialize_distributed(seed=0, heap_bytes=1 << 30)
W, me = U.world_size(), U.rank()
dev = U.current_device(); grp = U.get_triton_dist_world()
cfg = ModelConfig(model_name=args.model, max_length=args.ctx + 64, dtype=torch.bfloat16, rank=me, world_size=W,
                  num_layers_override=args.layers or None)
m = AutoLLM.from_pretrained(cfg, grp)
B = args.bs
def new_kv():
    kv = KV_Cache(m.num_layers, B, cfg.max_length, m.num_key_value_heads, m.head_dim, torch.bfloat16, W, dev)
    kv.rand_fill_kv_cache(args.ctx)
    return kv
ids = torch.randint(0, 1000, (B, 1), device=dev)
res = {}
def timed(fn, steps=args.steps, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

def graphed(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    torch.cuda.synchronize(); dist.barrier()
    return g.replay

for backend in ("torch", "triton_dist_AR", "triton_dist_gemm_ar"):
    kv = new_kv()
    pos = kv.kv_offset.to(torch.int64)[:, None].contiguous()
    m.set_fwd(backend)
    if backend == "triton_dist_AR": m.init_triton_dist_AR_ctx(max_M=B)
    if backend == "triton_dist_gemm_ar": m.init_triton_dist_gemm_ar_ctx(max_M=B)
    U.barrier_all_host()
    step = lambda: m.inference(ids, pos, kv)
    if backend == "torch": res["torch_eager_ms"] = timed(step, 10, 3)
    try:
        res[f"{backend}_graph_ms"] = timed(graphed(step))
    except Exception as e:
        res[f"{backend}_graph_ms"] = f"error: {str(e)[:120]}"
    m.finalize(); del kv
kv = new_kv()
mega = MegaDenseModel(m, B, kv)
stepm = lambda: mega.mega_forward(ids)
res["megakernel_ms"] = timed(stepm)
try:
    res["megakernel_graph_ms"] = timed(graphed(stepm))
except Exception as e:
    res["megakernel_graph_ms"] = f"error: {str(e)[:120]}"
res.update(model=args.model, tp=W, bs=B, ctx=args.ctx, layers=m.num_layers, sm_activity=mega.builder.get_sm_activity())
for k in list(res):
    if k.endswith("_ms") and isinstance(res[k], float): res[k.replace("_ms", "_tok_s")] = round(B * 1e3 / res[k], 1)
if me == 0:
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open(f"gpurun_out/qwen3_decode_tp{W}.json", "w"), indent=1)
U.finalize_distributed()
