"""Tutorial 11 -- communication kernels written in Python: the OpenSHMEM-style device API, scoped memory operations and a token-routing
protocol, all in the ``triton_dist.lk`` DSL.

Reference material this corresponds to: tutorials/01 (notify / wait), the ``libshmem_device`` / ``language_extra`` vocabularies that
reference kernels are written in, tutorials/04 (DeepSeek-style all-to-all) and little_kernel's FlashComm EP port.

Every kernel below is ONE Python function.  On GPUs it is compiled by nvcc for sm_100a; without GPUs the same source runs in the CPU
interpreter and the ranks talk through the shared-memory emulation heap:
    bash scripts/launch.sh --nproc_per_node=2 tutorials/11_distributed_kernels_in_python.py
    TD_FORCE_HOST_BACKEND=1 bash scripts/launch.sh --nproc_per_node=2 tutorials/11_distributed_kernels_in_python.py
"""
import torch
import triton_dist.utils as U
from triton_dist import lk
from triton_dist.lk import language_extra as le
from triton_dist.lk import ll, shmem

U.initialize_distributed(seed=0)
W, me, dev = U.world_size(), U.rank(), U.current_device()
gpu = dev.type == "cuda"


def run(kernel, grid, *args):
    if gpu:
        kernel[grid](*args)
        torch.cuda.synchronize()
    else:
        kernel.interpret(grid, *args)


# ---- 1. put-with-signal around a ring, then a barrier: the shmem vocabulary -------------------------------------------------------
# The first argument of everything that touches the symmetric heap is the kernel's ``ll.SymmCtx`` (rank, world, heap base / stride).
# ``shmem.make_sync`` wraps the two words the barriers advance: symmetric slots (uint32[2 * world], zeroed once) and a local epoch.
@lk.kernel(block=128)
def ring_exchange(ctx: ll.SymmCtx, slots: ll.ptr[ll.u32], epoch: ll.ptr[ll.u32], inbox: ll.ptr[ll.f32], sig: ll.ptr[ll.u64],
                  src: ll.ptr[ll.f32], n: ll.i32, phase: ll.u64):
    nxt = (shmem.my_pe(ctx) + 1) % shmem.n_pes(ctx)
    # the whole block copies (16-byte vectors over NVLink), then ONE thread stores the signal with release semantics
    shmem.putmem_signal_block(ctx, inbox, src, n * 4, sig, phase, shmem.SIGNAL_SET, nxt)
    if ll.threadIdx.x == 0:
        shmem.signal_wait_until(sig, shmem.CMP_GE, phase)         # acquire: my predecessor's data is visible after this
    shmem.barrier_all_block(ctx, shmem.make_sync(slots, epoch))    # nobody leaves before everybody has received


n = 256
inbox = U.nvshmem_create_tensor((n,), torch.float32)
sig = U.nvshmem_create_tensor((2,), torch.int64)
slots = U.nvshmem_create_tensor((2 * W,), torch.int32)
epoch = torch.zeros(1, dtype=torch.int32, device=dev)
for t in (inbox, sig, slots):
    t.zero_()
U.barrier_all_on_stream()
ctx = lk.symm_ctx()
for phase in (1, 2, 3):                                            # phase-numbered signals: no reset between calls
    src = torch.full((n,), 10.0 * me + phase, device=dev)
    run(ring_exchange, 1, ctx, slots, epoch, inbox, sig, src, n, phase)
    assert torch.equal(inbox.cpu(), torch.full((n,), 10.0 * ((me - 1) % W) + phase))
U.dist_print(f"rank {me}: ring put-with-signal x3 OK", allowed_ranks=[0])


# ---- 2. the memory model spelled out: scope= / semantic= on loads, stores and atomics ------------------------------------------------
# ``le.st(p, v, scope="sys", semantic="release")`` is one ``st.release.sys`` instruction, ``le.ld(..., "acquire")`` one ``ld.acquire.sys``.
# A counter on rank 0 collects one ticket per rank (remote atomic through the peer mapping); rank 0 publishes the total.
@lk.kernel(block=32)
def tickets(ctx: ll.SymmCtx, counter: ll.ptr[ll.u32], total: ll.ptr[ll.u32], my_ticket: ll.ptr[ll.u32]):
    if le.tid(0) == 0:
        c0 = ll.symm_at(ctx, counter, 0)                           # rank 0's copy of `counter`
        my_ticket[0] = le.atomic_add(c0, 1, scope="sys", semantic="acq_rel")
        if ll.rank(ctx) == 0:
            while le.ld(counter, scope="sys", semantic="acquire") < ll.u32(ll.num_ranks(ctx)):
                pass
            for r in range(ll.num_ranks(ctx)):
                le.st(ll.symm_at(ctx, total, r), le.ld(counter, scope="sys", semantic="relaxed"), scope="sys", semantic="release")
        le.wait_eq(total, ll.u32(ll.num_ranks(ctx)))


counter = U.nvshmem_create_tensor((1,), torch.int32)
total = U.nvshmem_create_tensor((1,), torch.int32)
for t in (counter, total):
    t.zero_()
U.barrier_all_on_stream()
ticket = torch.zeros(1, dtype=torch.int32, device=dev)
run(tickets, 1, ctx, counter, total, ticket)
seen = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(W)]
torch.distributed.all_gather(seen, ticket)
assert sorted(int(t.item()) for t in seen) == list(range(W)) and int(total.item()) == W
U.dist_print(f"rank {me}: {W} distinct tickets from one remote atomic counter OK", allowed_ranks=[0])


# ---- 3. expert-parallel dispatch / combine (token routing) as two DSL kernels ------------------------------------------------------
# triton_dist/lk/kernels/ep_a2a.py: a warp per (token, k) pair, slots from warp-aggregated atomics on LOCAL counters (every source owns a
# region on every destination), rows + tags stored to the peer, counts and phase-numbered flags published after a grid barrier;
# combine pulls the processed rows back and accumulates in fp32.
from triton_dist.lk.kernels.ep_a2a import LkEpAllToAll  # noqa: E402

T, H, topk, epr = (64, 256, 2, 2) if gpu else (8, 32, 2, 2)
E = W * epr
ep = LkEpAllToAll(T, H, topk, E)
g = torch.Generator().manual_seed(me)
x = (torch.randn(T, H, generator=g) * 0.5).to(torch.bfloat16).to(dev)
ids = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(T)]).to(torch.int32).to(dev)
w = torch.rand(T, topk, generator=g).to(dev)
recv_x, meta, cnt = ep.dispatch(x, ids)
if gpu:
    torch.cuda.synchronize()
for src in range(W):                                               # "experts": scale every received row by (global expert id + 1)
    k = int(cnt[src])
    if k:
        scale = (meta[src, :k, 1] + me * epr + 1).float()[:, None]
        ep.y_buf[src, :k] = (recv_x[src, :k].float() * scale).to(torch.bfloat16)
out = ep.combine(w)
ref = sum((x.float() * (ids[:, k] + 1).float()[:, None]).to(torch.bfloat16).float() * w[:, k, None] for k in range(topk))
torch.testing.assert_close(out.float().cpu(), ref.cpu(), atol=3e-2, rtol=3e-2)
U.dist_print(f"rank {me}: dispatch -> experts -> combine ({int(cnt.sum())} rows received) OK", allowed_ranks=[0])

# ---- 4. the NVSwitch as a reduction engine: multimem loads / stores from Python ------------------------------------------------------
# ``ll.symm_mc(ctx, p)`` is the multicast alias of a symmetric pointer: ``ll.multimem_ld_reduce_*`` through it returns the SUM of the word
# over every rank's copy (added inside the switch), ``ll.multimem_st_v4`` writes every copy.  triton_dist/lk/kernels/allreduce_nvls.py
# builds the one-shot and two-shot all-reduce of the product from them; the interpreter models the alias on the emulation heap.
from triton_dist.lk.kernels.allreduce_nvls import LkNvlsAllReduce  # noqa: E402

if not gpu or U.is_nvshmem_multimem_supported():
    for method in ("one_shot", "two_shot"):
        ar = LkNvlsAllReduce(4096, method)
        xs = torch.full((256,), float(me + 1), device=dev).bfloat16()
        got = ar(xs)
        assert torch.equal(got.float().cpu(), torch.full((256,), W * (W + 1) / 2.0))
        U.barrier_all_on_stream()
        ar.finalize()
    U.dist_print(f"rank {me}: NVLS one-shot / two-shot all-reduce OK", allowed_ranks=[0])

ep.finalize()
for t in (total, counter, slots, sig, inbox):
    U.nvshmem_free_tensor_sync(t)
U.dist_print("distributed DSL kernels tutorial OK", allowed_ranks=[0])
U.finalize_distributed()
