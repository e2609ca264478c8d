"""Tutorial 02 -- fast AllGather / AllReduce kernels (reference: tutorials/02,05,06; kernels/nvidia/allreduce.py,
low_latency_allgather.py).  One-shot / two-shot, P2P and NVLS (multimem) variants; per-CTA flag barriers; no resets."""
import torch
import torch.distributed as dist
import triton_dist.utils as U
from triton_dist.ops import comm

U.initialize_distributed()
W, me = U.world_size(), U.rank()
dev = U.current_device()
ag = comm.create_fast_allgather_context(1 << 20)
ar = comm.create_allreduce_ctx(1 << 20, me, W, W)
x = torch.randn(4096, device=dev).to(torch.bfloat16 if dev.type == "cuda" else torch.float32)
for mode in (("pull", "push", "push_2d_ll") if dev.type == "cuda" else ("push",)):
    out = comm.fast_allgather(x, ag, mode=mode)
    ref = torch.empty(W * x.numel(), dtype=x.dtype, device=dev)
    dist.all_gather_into_tensor(ref, x)
    assert torch.equal(out.view(-1), ref)
methods = comm.get_allreduce_methods() if dev.type == "cuda" else [comm.AllReduceMethod.OneShot]
for m in methods:
    y = comm.all_reduce(x, m, ar)
    ref = x.clone(); dist.all_reduce(ref)
    torch.testing.assert_close(y.float(), ref.float(), atol=0.1, rtol=2e-2)
U.dist_print(f"multimem={U.is_nvshmem_multimem_supported()} auto(64KB)={comm.get_auto_allreduce_method(65536).name} OK", allowed_ranks=[0])
ag.finalize(); ar.finalize(); U.finalize_distributed()
