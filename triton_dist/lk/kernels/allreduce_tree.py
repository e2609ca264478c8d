"""Double-binary-tree all-reduce over the symmetric heap, in the DSL.

Reference: kernels/nvidia/allreduce.py ``allreduce_double_tree_intra_node_kernel`` (:216-330: an up pass and a down pass over two
complementary binary trees, per-chunk signals).  On one NVSwitch domain the product's all-reduce is the NVLS one-/two-shot pair
(``AllReduceMethod.DoubleTree`` selects the two-shot kernel there); this kernel is the tree algorithm itself -- log2(W) hops instead of
W - 1, every rank interior in at most one tree so both directions of every link carry payload -- for fabrics without in-switch reduction,
and an example of a two-phase protocol written in Python:

* the message is split in two halves; tree A (rooted at rank 0) reduces the first, tree B (rooted at rank W - 1) the second
  (``triton_dist.ops.comm.get_tree_parent_and_children``);
* up pass: a rank waits for its children's partials (pushed into its ``up`` slots), adds its own half and pushes the sum into its parent's
  slot (left or right child); the root now holds the result;
* down pass: the root writes ``out`` and pushes the result into its children's ``down`` buffer, every other rank waits for its parent's
  push, writes ``out`` and forwards;
* every CTA owns a slice of each half and runs both passes for it on its own flags; flags carry the call number, nothing is reset.  No
  double buffering is needed: a child can only start the next call after its parent's down push, which the parent issues after it has
  consumed the child's up partial; a parent can only overwrite a child's ``down`` buffer after the child's next up push, which the child
  issues after it has finished reading the previous one.

fp32 payload.  ``tests/dist_worker.py`` case ``lk_ar_tree`` runs it across processes in the interpreter (world 2 and 4, chaos) against
``torch.distributed.all_reduce``.
"""
from triton_dist import lk
from triton_dist.lk import ll

THREADS = 256


@lk.kernel(block=THREADS)
def allreduce_double_tree(ctx: ll.SymmCtx, x: ll.ptr[ll.f32], out: ll.ptr[ll.f32], up: ll.ptr[ll.f32], down: ll.ptr[ll.f32],
                          flags: ll.ptr[ll.u32], half: ll.i32, max_half: ll.i32, rel: ll.ptr[ll.i32], phase: ll.u32):
    """``rel``: int32 [8] = (parent, left, right, my slot in the parent) of tree A, then of tree B; -1 = none."""
    G = ll.gridDim.x
    cta = ll.blockIdx.x
    tid = ll.threadIdx.x
    per = (half + G - 1) // G
    lo = cta * per
    hi = min(half, lo + per)
    for tree in ll.static_range(2):
        parent = rel[tree * 4 + 0]
        left = rel[tree * 4 + 1]
        right = rel[tree * 4 + 2]
        slot = rel[tree * 4 + 3]
        base = ll.i64(tree) * half                                            # my half of x / out
        up0 = up + ll.i64(tree * 2) * max_half                                # my two child slots of this tree
        dn = down + ll.i64(tree) * max_half
        f_up = flags + (tree * 2) * G
        f_dn = flags + (4 + tree) * G
        # ---- up pass ----
        if tid < 32:
            if left >= 0:
                ll.wait(f_up + cta, 1, phase, True)
            if right >= 0:
                ll.wait(f_up + (G + cta), 1, phase, True)
        ll.syncthreads()
        if parent >= 0:
            dst = ll.symm_at(ctx, up, parent) + (ll.i64(tree * 2 + slot) * max_half)
        else:
            dst = out + base
        for i in range(lo + tid, hi, THREADS):
            v = x[base + i]
            if left >= 0:
                v += up0[i]
            if right >= 0:
                v += up0[max_half + i]
            dst[i] = v
            if parent < 0:                                                    # root: the sum is final -- start the down pass
                if left >= 0:
                    cl = ll.symm_at(ctx, down, left) + ll.i64(tree) * max_half
                    cl[i] = v
                if right >= 0:
                    cr = ll.symm_at(ctx, down, right) + ll.i64(tree) * max_half
                    cr[i] = v
        ll.syncthreads()
        if tid == 0:
            if parent >= 0:
                ll.notify(ctx, f_up + (slot * G + cta), parent, phase)
            else:
                if left >= 0:
                    ll.notify(ctx, f_dn + cta, left, phase)
                if right >= 0:
                    ll.notify(ctx, f_dn + cta, right, phase)
        # ---- down pass ----
        if parent >= 0:
            if tid < 32:
                ll.wait(f_dn + cta, 1, phase, True)
            ll.syncthreads()
            for i2 in range(lo + tid, hi, THREADS):
                r = dn[i2]
                out[base + i2] = r
                if left >= 0:
                    dl = ll.symm_at(ctx, down, left) + ll.i64(tree) * max_half
                    dl[i2] = r
                if right >= 0:
                    dr = ll.symm_at(ctx, down, right) + ll.i64(tree) * max_half
                    dr[i2] = r
            ll.syncthreads()
            if tid == 0:
                if left >= 0:
                    ll.notify(ctx, f_dn + cta, left, phase)
                if right >= 0:
                    ll.notify(ctx, f_dn + cta, right, phase)


class LkDoubleTreeAllReduce:
    """Symmetric ``up`` [2 trees, 2 child slots, max_half] / ``down`` [2 trees, max_half] buffers and flags for messages of up to
    ``2 * max_half`` fp32 elements; the world size must be a power of two."""

    def __init__(self, max_elems: int, grid: int = 0):
        import torch
        import triton_dist.utils as U
        from triton_dist.ops.comm import get_tree_parent_and_children
        self.W, self.rank = U.world_size(), U.rank()
        assert self.W & (self.W - 1) == 0, "the double binary tree needs a power-of-two world size"
        self.max_half = (max_elems + 1) // 2
        dev = U.current_device()
        self.gpu = dev.type == "cuda"
        self.grid = grid or (min(32, max(1, self.max_half // 4096)) if self.gpu else 2)
        self.up = U.nvshmem_create_tensor((4 * self.max_half,), torch.float32)
        self.down = U.nvshmem_create_tensor((2 * self.max_half,), torch.float32)
        self.flags = U.nvshmem_create_tensor((6 * self.grid,), torch.int32)
        self.flags.zero_()
        rel = []
        if self.W > 1:
            t = get_tree_parent_and_children(self.W, self.rank)
            for (p, l, r) in (t[:3], t[3:]):
                slot = -1
                if p >= 0:
                    pt = get_tree_parent_and_children(self.W, p)
                    kids = pt[1:3] if (p, l, r) == t[:3] else pt[4:6]
                    slot = 0 if kids[0] == self.rank else 1
                rel += [p, l, r, slot]
        else:
            rel = [-1, -1, -1, -1] * 2
        self.rel = torch.tensor(rel, dtype=torch.int32, device=dev)
        self.phase = 0
        U.barrier_all_on_stream()

    def __call__(self, x, out=None):
        """x: fp32, even number of elements <= 2 * max_half -> the sum over all ranks (same shape)."""
        import torch
        n = x.numel()
        assert x.dtype == torch.float32 and n % 2 == 0 and n // 2 <= self.max_half
        out = torch.empty_like(x) if out is None else out
        self.phase += 1
        args = (lk.symm_ctx(), x.contiguous().view(-1), out.view(-1), self.up, self.down, self.flags, n // 2, self.max_half, self.rel, self.phase)
        if self.gpu:
            allreduce_double_tree[self.grid](*args)
        else:
            allreduce_double_tree.interpret(self.grid, *args)
        return out

    def finalize(self):
        import triton_dist.utils as U
        for t in (self.flags, self.down, self.up):
            U.nvshmem_free_tensor_sync(t)
