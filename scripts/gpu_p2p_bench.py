"""NVLink mechanism micro-benchmark (2+ GPUs): GB/s of push/pull via generic ld/st vs TMA bulk, vs #CTAs.
Also measures copy-engine peer copy (torch copy_) for reference."""
import ctypes as C, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import triton_dist.utils as U
from triton_dist import _C
_C.register("td_p2p_bench", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p])
U.initialize_distributed(seed=0, heap_bytes=1 << 30)
W, me = U.world_size(), U.rank()
lib = _C.cuda_lib()
nbytes = 256 << 20
buf = U.nvshmem_create_tensor((nbytes,), torch.uint8)
local = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
peer = U.symm_at(buf, (me + 1) % W)
rows = []
def run(mode, grid, threads, dst, src, label):
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2): lib.td_p2p_bench(mode, grid, threads, dst.data_ptr(), src.data_ptr(), nbytes, st)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lib.td_p2p_bench(mode, grid, threads, dst.data_ptr(), src.data_ptr(), nbytes, st)
    e1.record(); torch.cuda.synchronize()
    gbs = 5 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    r = dict(label=label, mode=mode, grid=grid, threads=threads, gbs=round(gbs, 1), gbs_per_sm=round(gbs / grid, 2))
    rows.append(r)
    if me == 0: print(json.dumps(r), flush=True)
for grid in (4, 16, 32, 64, 148):
    for threads in (256, 1024):
        run(0, grid, threads, peer, local, "generic push (all ranks push to rank+1)")
        run(1, grid, threads, local, peer, "generic pull")
    run(2, grid, 32, peer, local, "TMA bulk push")
    run(3, grid, 32, local, peer, "TMA bulk pull")
    run(2, grid, 32, buf, local, "TMA bulk local copy")
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): peer.copy_(local)
e1.record(); torch.cuda.synchronize()
r = dict(label="copy engine push (torch copy_)", gbs=round(5 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1))
rows.append(r)
if me == 0:
    print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True); json.dump(rows, open(f"gpurun_out/p2p_bench_n{W}.json", "w"), indent=1)
U.finalize_distributed()
