"""Tutorial 01 -- notify / wait on the symmetric heap (reference: tutorials/01-distributed-notify-wait.py).

Each rank writes a payload into its right neighbour's queue slot, then raises a flag there with release semantics;
the neighbour waits for the flag with acquire semantics and reads the payload.  Flags carry the round number, so
nothing is ever reset.  Runs on GPUs (NVLink P2P) or without any GPU (shared-memory emulation):
    bash scripts/launch.sh --nproc_per_node=2 tutorials/01_notify_wait.py
    TD_FORCE_HOST_BACKEND=1 bash scripts/launch.sh --nproc_per_node=2 tutorials/01_notify_wait.py
The device-side equivalents used inside the fused kernels are td::notify / td::wait / td::symm_at (csrc/td/primitives.cuh).
"""
import torch
import triton_dist.utils as U
from triton_dist import language as dl

U.initialize_distributed()
W, me = U.world_size(), U.rank()
queue = U.nvshmem_create_tensor((1024,), torch.float32)
signal = U.nvshmem_create_tensor((8,), torch.int32)
U.barrier_all_on_stream()
nxt, prv = (me + 1) % W, (me - 1 + W) % W
for rnd in range(1, 21):
    data = torch.full((1024,), float(me + rnd * 100), device=queue.device)
    dl.symm_at(queue, nxt).copy_(data)                             # producer: remote store into the peer's segment
    dl.notify(signal[0:1], nxt, signal=rnd, sig_op="set")          # release
    tok = dl.wait(signal[0:1], 1, "sys", "acquire", wait_value=rnd)  # consumer: acquire
    got = dl.consume_token(queue, tok).clone()
    assert torch.all(got == prv + rnd * 100)
    U.barrier_all_on_stream()
U.dist_print(f"rank {me}: 20 rounds OK", need_sync=True)
U.finalize_distributed()
