"""MegaKernel: the whole decode step of a tensor-parallel dense LLM as ONE persistent CUDA kernel.

Reference: /root/reference/python/triton_dist/mega_triton_kernel/** (``ModelBuilder.make_*`` -> task graph ->
scheduler -> generated Triton kernel -> ``run()``; ``DenseModel.mega_forwrad``).  Here: ``ModelBuilder`` records ops
and tiles them into tasks with scoreboard dependencies, ``schedule()`` deals them to per-CTA queues (round-robin or
zig-zag, as core/scheduler.py:103-168), and the interpreter is the hand-written kernel in csrc/megakernel.cu.
On the emulation backend ``run()`` interprets the same task list with PyTorch ops (same order, same buffers).
"""
from __future__ import annotations

import ctypes as C
import enum
import math
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from .. import _C
from .. import utils as U
from ..ops.comm import SymmArgs, symm_args

T_RMSNORM, T_LINEAR, T_QKROPE, T_ATTN, T_ALLREDUCE, T_ATTN_COMBINE, T_SILU_MUL, T_ADD, T_PREFETCH = 1, 2, 3, 4, 5, 7, 8, 9, 10
T_QKROPE_PAGED, T_ATTN_PAGED = 11, 12      # paged KV cache (block table; reference: mega_triton_kernel/models/paged_kv_cache.py)
T_FLASH_ATTN, T_QKROPE_SPLIT = 13, 14      # prefill: tensor-core attention over [B, S, H, 128] and the qk-norm + rope that feeds it
FLASH_BQ = 128                              # query rows per FLASH_ATTN task (8 warps x 16)
T_JOIN = 15                                 # no-op: waits on its dependency, signals its counter (several producers -> one counter)
TASK_NAMES = {1: "rms_norm", 2: "linear", 3: "qk_norm_rope_update_kvcache", 4: "flash_decode", 5: "allreduce", 7: "flash_decode_combine", 8: "silu_mul_up", 9: "add", 10: "prefetch", 11: "qk_norm_rope_update_paged_kvcache", 12: "flash_decode_paged", 13: "flash_attn", 14: "qkv_pack_qk_norm_rope_split_v", 15: "join"}


class _MegaArgs(C.Structure):
    _fields_ = [("symm", SymmArgs), ("tasks", C.c_void_p), ("queue_off", C.c_void_p), ("ptrs", C.c_void_p), ("sb", C.c_void_p),
                ("epoch", C.c_void_p), ("B", C.c_longlong), ("grid", C.c_longlong), ("smem_bytes", C.c_longlong),
                ("dynamic", C.c_longlong), ("num_tasks", C.c_longlong)]


_C.register("td_mega_launch", C.c_int, [C.POINTER(_MegaArgs), C.c_void_p])
_C.register("td_mega_task_size", C.c_int, [])


def _fbits(x: float) -> int:
    return struct.unpack("i", struct.pack("f", x))[0]


@dataclass
class Task:
    type: int
    dep_idx: int
    dep_count: int
    sig_idx: int
    args: List[int]
    layer: int = 0

    def pack(self) -> List[int]:
        a = list(self.args) + [0] * (12 - len(self.args))
        return [self.type, self.dep_idx, self.dep_count, self.sig_idx] + a


class SchedulingStrategy(enum.Enum):
    """Reference: mega_triton_kernel/core/scheduler.py ``SchedulingStrategy``."""
    ROUND_ROBIN = "round_robin"
    ZIG_ZAG = "zig_zag"
    RUNTIME = "dynamic"


def round_robin_scheduler(tasks: Sequence, num_queues: int) -> List[list]:
    """Task i -> queue i % n: consecutive tiles of one op land on different SMs."""
    queues: List[list] = [[] for _ in range(num_queues)]
    for i, t in enumerate(tasks):
        queues[i % num_queues].append(t)
    return queues


def zig_zag_scheduler(tasks: Sequence, num_queues: int) -> List[list]:
    """Alternate the direction of every sweep (0..n-1, n-1..0, ...): the SM that got the last tile of a sweep -- typically the one that
    finishes last -- gets the first tile of the next, which evens out per-SM load when tile counts are not multiples of n."""
    queues: List[list] = [[] for _ in range(num_queues)]
    for i, t in enumerate(tasks):
        sweep, pos = divmod(i, num_queues)
        queues[pos if sweep % 2 == 0 else num_queues - 1 - pos].append(t)
    return queues


def enque_tasks(tasks: Sequence, num_queues: int, strategy="round_robin") -> List[list]:
    """Tasks (in program = topological order) -> per-CTA work queues.  ``dynamic``: queue 0 holds everything, the kernel's CTAs pop from
    it with an atomic counter."""
    name = strategy.value if isinstance(strategy, SchedulingStrategy) else str(strategy)
    if name == "dynamic":
        return [list(tasks)] + [[] for _ in range(num_queues - 1)]
    if name == "zig_zag":
        return zig_zag_scheduler(tasks, num_queues)
    if name == "round_robin":
        return round_robin_scheduler(tasks, num_queues)
    raise ValueError(f"unknown scheduling strategy {strategy!r}")


def work_queue_list_to_device_tensor(queues: Sequence[Sequence["Task"]], device=None):
    """(task tensor int32 [num_tasks, 16] in queue order, queue offsets int32 [num_queues + 1]) -- what the kernel indexes."""
    flat, off = [], [0]
    for q in queues:
        for t in q:
            flat.extend(t.pack())
        off.append(off[-1] + len(q))
    return (torch.tensor(flat, dtype=torch.int32, device=device).view(-1, 16) if flat else torch.zeros((0, 16), dtype=torch.int32, device=device),
            torch.tensor(off, dtype=torch.int32, device=device))


class ModelBuilder:
    """Records ops of one decode step, tiles them into tasks, schedules them, launches the persistent kernel."""

    def __init__(self, batch: int, num_sms: Optional[int] = None, schedule: str = "round_robin", auto_deps: bool = False):
        assert 1 <= batch <= 64, "megakernel decode batch: 1..8 (GEMV tasks) or 9..64 (tensor-core tasks)"
        self.B = batch
        self.device = U.current_device()
        self.is_cuda = self.device.type == "cuda"
        self.num_sms = num_sms or (torch.cuda.get_device_properties(self.device).multi_processor_count if self.is_cuda else 8)
        self.schedule_policy = schedule
        # auto_deps: ops called WITHOUT ``dep=`` wait for whatever earlier ops wrote their inputs or still read / write their outputs
        # (the reference derives tile dependencies from a buffer graph, mega_triton_kernel/core/graph.py; here: interval overlap of
        # the tensors' storage, one JOIN task per extra producer when there are several)
        self.auto_deps = auto_deps
        self._writers: List[tuple] = []         # (lo, hi, (counter, count)) of the op that last wrote the bytes
        self._readers: List[tuple] = []         # ops that read the bytes since
        self.tasks: List[Task] = []
        self.ptrs: List[torch.Tensor] = []
        self._ptr_idx: Dict[int, int] = {}
        self.n_counters = 0
        self.max_smem = 8 * 8 * 130 * 4 + 1024
        self.compiled = False
        self.cur_layer = 0
        self.metrics: Dict[str, int] = {}

    # ---- bookkeeping ----
    def ptr(self, t: Optional[torch.Tensor]) -> int:
        if t is None:
            return -1
        key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()))   # views that start at the same address are distinct entries (host interpretation)
        if key not in self._ptr_idx:
            self._ptr_idx[key] = len(self.ptrs)
            self.ptrs.append(t)
        return self._ptr_idx[key]

    def ptr_pair(self, t: torch.Tensor, follower: torch.Tensor) -> int:
        """Two ADJACENT pointer-table entries (t, follower): the paged tasks find the block table at ``index of the value cache + 1``
        without spending another task argument."""
        idx = len(self.ptrs)
        self.ptrs += [t, follower]
        return idx

    def counter(self) -> int:
        self.n_counters += 1
        return self.n_counters - 1

    def _add(self, t: Task):
        t.layer = self.cur_layer
        self.tasks.append(t)
        self.metrics[TASK_NAMES[t.type]] = self.metrics.get(TASK_NAMES[t.type], 0) + 1

    # ---- dependency inference ----
    @staticmethod
    def _span(t: torch.Tensor):
        es = t.element_size()
        n = sum((sz - 1) * st for sz, st in zip(t.shape, t.stride())) + 1 if t.numel() else 0
        return t.data_ptr(), t.data_ptr() + n * es

    def _infer_dep(self, reads, writes):
        producers = []

        def hit(entries, lo, hi):
            for (l, h, d) in entries:
                if l < hi and lo < h and d not in producers:
                    producers.append(d)
        for t in reads:
            if t is not None:
                hit(self._writers, *self._span(t))                      # read after write
        for t in writes:
            if t is not None:
                lo, hi = self._span(t)
                hit(self._writers, lo, hi)                               # write after write
                hit(self._readers, lo, hi)                               # write after read
        if not producers:
            return (-1, 0)
        if len(producers) == 1:
            return producers[0]
        j = self.counter()
        for (sig, cnt) in producers:
            self._add(Task(T_JOIN, sig, cnt, j, []))
        return (j, len(producers))

    def _record(self, result, reads, writes):
        if result is None:
            return
        for t in writes:
            if t is not None:
                lo, hi = self._span(t)
                self._writers = [(l, h, d) for (l, h, d) in self._writers if not (lo <= l and h <= hi)] + [(lo, hi, result)]
                self._readers = [(l, h, d) for (l, h, d) in self._readers if not (l < hi and lo < h)]
        for t in reads:
            if t is not None:
                self._readers.append((*self._span(t), result))

    # ---- ops (each returns (counter, count) that consumers wait on) ----
    def make_rms_norm(self, x, weight, out, eps: float, dep=None, residual=None, residual_out=None):
        sig = self.counter()
        d = dep or (-1, 0)
        self._add(Task(T_RMSNORM, d[0], d[1], sig, [self.ptr(x), self.ptr(residual), self.ptr(weight), self.ptr(out),
                                                     self.ptr(residual_out), x.shape[-1], _fbits(eps)]))
        return sig, 1

    def make_linear(self, x, weight, out, dep, act_silu_mul: bool = False, tile_n: Optional[int] = None, norm_weight=None,
                    eps: float = 1e-6):
        """out[B, N] = act(x) @ weight[N, K]^T (GEMV tiles).  ``act_silu_mul``: x is [B, 2K] = (gate | up).
        ``norm_weight``: RMSNorm(x) * norm_weight is applied while the operand is staged (no separate norm task)."""
        N, K = weight.shape
        tn = tile_n or max(8, ((N + self.num_sms - 1) // self.num_sms + 7) // 8 * 8)
        if self.B > 8:
            # tensor-core tasks (csrc/megakernel.cu linear_mma): 8-column groups; a power-of-two number of groups per tile lets the
            # warps of a CTA share a group along K when there are fewer than 8 of them
            assert K % 32 == 0 and N % 8 == 0, "tensor-core linear tasks need K % 32 == 0 and N % 8 == 0"
            if tile_n is None:
                tn = 8
                while tn < 128 and (N + tn - 1) // tn > self.num_sms:
                    tn *= 2
            assert tn % 8 == 0 and tn <= 128
        sig = self.counter()
        n_tiles = 0
        d = dep or (-1, 0)
        act = 2 if norm_weight is not None else int(act_silu_mul)
        assert not (norm_weight is not None and act_silu_mul)
        for n0 in range(0, N, tn):
            self._add(Task(T_LINEAR, d[0], d[1], sig, [self.ptr(x), self.ptr(weight), self.ptr(out), K, out.shape[-1], n0,
                                                       min(tn, N - n0), act, x.shape[-1], self.ptr(norm_weight), _fbits(eps)]))
            n_tiles += 1
        self.max_smem = max(self.max_smem, (128 * 1024 + 512) if self.B > 8 else (self.B * K * 2 + 256))
        return sig, n_tiles

    make_qkv_proj = make_o_proj = make_fc1 = make_fc2 = make_linear

    @staticmethod
    def _paged_arg(k_cache, block_table):
        """(page_size | max_pages << 16) for caches shaped [pages, page_size, Hkv, D] with a [B, max_pages] int32 block table."""
        page_size, max_pages = k_cache.shape[1], block_table.shape[1]
        assert block_table.dtype == torch.int32 and block_table.is_contiguous() and page_size < (1 << 16) and max_pages < (1 << 15)
        return page_size | (max_pages << 16)

    def make_qk_norm_rope_update_kvcache(self, qkv, q_out, k_cache, v_cache, q_norm_w, k_norm_w, positions, Hq, Hkv, eps, theta, dep,
                                         block_table=None):
        """``block_table`` (int32 [B, max_pages]): ``k_cache`` / ``v_cache`` are page pools [pages, page_size, Hkv, D] and the new token of
        sequence b is stored in page ``block_table[b, pos // page_size]`` (reference: the megakernel's paged KV cache)."""
        sig = self.counter()
        if block_table is None:
            self._add(Task(T_QKROPE, dep[0], dep[1], sig, [self.ptr(qkv), self.ptr(q_out), self.ptr(k_cache), self.ptr(v_cache),
                                                           self.ptr(q_norm_w), self.ptr(k_norm_w), self.ptr(positions), Hq, Hkv,
                                                           k_cache.shape[1], _fbits(eps), _fbits(theta)]))
        else:
            self._add(Task(T_QKROPE_PAGED, dep[0], dep[1], sig, [self.ptr(qkv), self.ptr(q_out), self.ptr(k_cache), self.ptr_pair(v_cache, block_table),
                                                                 self.ptr(q_norm_w), self.ptr(k_norm_w), self.ptr(positions), Hq, Hkv,
                                                                 self._paged_arg(k_cache, block_table), _fbits(eps), _fbits(theta)]))
        return sig, 1

    def make_flash_decode(self, q, k_cache, v_cache, positions, out, Hq, Hkv, sm_scale, dep, n_splits: int = 1, scratch=None,
                          block_table=None):
        """GQA decode.  ``n_splits`` > 1: split-KV -- every (batch, kv head) becomes n_splits tasks that write (m, l, o) partials
        to ``scratch`` (fp32 [B, Hkv, n_splits, 8, 130]) plus one combine task, so a handful of heads still fills the SMs.
        ``block_table``: paged KV cache (see ``make_qk_norm_rope_update_kvcache``)."""
        if block_table is not None:
            return self._make_flash_decode_paged(q, k_cache, v_cache, positions, out, Hq, Hkv, sm_scale, dep, n_splits, scratch, block_table)
        sig = self.counter()
        n = 0
        if n_splits <= 1:
            for b in range(self.B):
                for kvh in range(Hkv):
                    self._add(Task(T_ATTN, dep[0], dep[1], sig, [self.ptr(q), self.ptr(k_cache), self.ptr(v_cache), self.ptr(positions),
                                                                 self.ptr(out), b, kvh, Hq, Hkv, k_cache.shape[1], _fbits(sm_scale), 0]))
                    n += 1
            return sig, n
        assert scratch is not None and scratch.dtype == torch.float32 and scratch.numel() >= self.B * Hkv * n_splits * 8 * 130
        for b in range(self.B):
            for kvh in range(Hkv):
                for s in range(n_splits):
                    self._add(Task(T_ATTN, dep[0], dep[1], sig, [self.ptr(q), self.ptr(k_cache), self.ptr(v_cache), self.ptr(positions),
                                                                 self.ptr(scratch), b, kvh, Hq, Hkv, k_cache.shape[1], _fbits(sm_scale),
                                                                 s | (n_splits << 16)]))
                    n += 1
        sig2 = self.counter()
        m = 0
        for b in range(self.B):
            for kvh in range(Hkv):
                self._add(Task(T_ATTN_COMBINE, sig, n, sig2, [self.ptr(scratch), self.ptr(out), b, kvh, Hq, Hkv, n_splits]))
                m += 1
        return sig2, m

    def _make_flash_decode_paged(self, q, k_cache, v_cache, positions, out, Hq, Hkv, sm_scale, dep, n_splits, scratch, block_table):
        packed = self._paged_arg(k_cache, block_table)
        vidx = self.ptr_pair(v_cache, block_table)
        sig = self.counter()
        n = 0
        if n_splits <= 1:
            for b in range(self.B):
                for kvh in range(Hkv):
                    self._add(Task(T_ATTN_PAGED, dep[0], dep[1], sig, [self.ptr(q), self.ptr(k_cache), vidx, self.ptr(positions),
                                                                       self.ptr(out), b, kvh, Hq, Hkv, packed, _fbits(sm_scale), 0]))
                    n += 1
            return sig, n
        assert scratch is not None and scratch.dtype == torch.float32 and scratch.numel() >= self.B * Hkv * n_splits * 8 * 130
        for b in range(self.B):
            for kvh in range(Hkv):
                for s in range(n_splits):
                    self._add(Task(T_ATTN_PAGED, dep[0], dep[1], sig, [self.ptr(q), self.ptr(k_cache), vidx, self.ptr(positions),
                                                                       self.ptr(scratch), b, kvh, Hq, Hkv, packed, _fbits(sm_scale),
                                                                       s | (n_splits << 16)]))
                    n += 1
        sig2 = self.counter()
        m = 0
        for b in range(self.B):
            for kvh in range(Hkv):
                self._add(Task(T_ATTN_COMBINE, sig, n, sig2, [self.ptr(scratch), self.ptr(out), b, kvh, Hq, Hkv, n_splits]))
                m += 1
        return sig2, m

    # ---- prefill (reference: model_builder.py make_flash_attn / make_qkv_pack_flash_attn / make_qkv_pack_qk_norm_rope_split_v) ----
    def make_flash_attn(self, q, k, v, output, sm_scale=None, soft_cap: float = 0.0, is_causal: bool = True, dep=None):
        """output[B, S, Hq, 128] = attention(q[B, S, Hq, 128], k / v[B, S, Hkv, 128]); one task per (batch, q head, 128 query rows) on
        the warp-level tensor cores.  q / k / v may be views into one packed qkv tensor (token stride a multiple of 128 elements)."""
        Bq, S, Hq, d = q.shape
        Hkv = k.shape[2]
        assert d == 128 and tuple(k.shape) == tuple(v.shape) == (Bq, S, Hkv, d) and Hq % Hkv == 0 and tuple(output.shape) == (Bq, S, Hq, d)
        assert output.is_contiguous() and k.stride(1) == v.stride(1)
        for t in (q, k, v, output):
            assert t.dtype == torch.bfloat16 or not self.is_cuda, "the FLASH_ATTN task is a bf16 kernel"
            assert t.stride(3) == 1 and t.stride(2) == d and t.stride(1) % d == 0 and t.stride(0) == S * t.stride(1)
        q_h, kv_h = q.stride(1) // d, k.stride(1) // d
        assert q_h < (1 << 16) and kv_h < (1 << 15) and Hq < (1 << 16) and Hkv < (1 << 15)
        scale = float(sm_scale) if sm_scale is not None else d ** -0.5
        self.max_smem = max(self.max_smem, (64 * 80 + 128 * 36) * 4)
        self.has_prefill = True
        sig = self.counter()
        dd = dep or (-1, 0)
        n = 0
        for b in range(Bq):
            for h in range(Hq):
                for qb in range((S + FLASH_BQ - 1) // FLASH_BQ):
                    self._add(Task(T_FLASH_ATTN, dd[0], dd[1], sig, [self.ptr(q), self.ptr(k), self.ptr(v), self.ptr(output), b, h,
                                                                     qb | (int(bool(is_causal)) << 30), S, Hq | (Hkv << 16), q_h | (kv_h << 16),
                                                                     _fbits(scale), _fbits(float(soft_cap))]))
                    n += 1
        return sig, n

    def make_qkv_pack_flash_attn(self, qkv, output, sm_scale=None, soft_cap: float = 0.0, is_causal: bool = True, dep=None):
        """qkv: [B, S, Hq + 2 Hkv, 128] packed, output: [B, S, Hq, 128]."""
        Hq = output.shape[2]
        Hkv = (qkv.shape[2] - Hq) // 2
        return self.make_flash_attn(qkv[:, :, :Hq], qkv[:, :, Hq:Hq + Hkv], qkv[:, :, Hq + Hkv:], output, sm_scale, soft_cap, is_causal, dep)

    def make_qkv_pack_qk_norm_rope_split_v(self, qkv, kv_lens, q_norm_w, k_norm_w, q_out, k_out, v_out, eps: float, theta: float, dep=None):
        """Prefill form of the qk-norm + rope task: qkv [B, S, Hq + 2 Hkv, 128] -> q_out [B, S, Hq, 128], k_out / v_out [B, S, Hkv, 128];
        token (b, s) is rotated at position ``kv_lens[b] + s``.  The builder's batch is the token count B * S."""
        Bq, S, heads, d = qkv.shape
        Hq, Hkv = q_out.shape[2], k_out.shape[2]
        assert d == 128 and heads == Hq + 2 * Hkv and Bq * S == self.B, "create the builder with batch = B * S tokens"
        assert qkv.is_contiguous() and q_out.is_contiguous() and k_out.is_contiguous() and v_out.is_contiguous() and kv_lens.dtype == torch.int32
        self.has_prefill = True
        sig = self.counter()
        dd = dep or (-1, 0)
        self._add(Task(T_QKROPE_SPLIT, dd[0], dd[1], sig, [self.ptr(qkv), self.ptr(q_out), self.ptr(k_out), self.ptr(v_out), self.ptr(q_norm_w),
                                                           self.ptr(k_norm_w), self.ptr(kv_lens), Hq, Hkv, S, _fbits(eps), _fbits(theta)]))
        return sig, 1

    def make_allreduce(self, part_symm, flags_symm, residual, residual_out, dep, n_slices: int = 4):
        """residual_out = residual + sum over ranks of part (one-shot over NVLink; fused residual add)."""
        nvec = part_symm.numel() // 8
        n_slices = max(1, min(n_slices, nvec))
        per = (nvec + n_slices - 1) // n_slices
        sig = self.counter()
        for s in range(n_slices):
            self._add(Task(T_ALLREDUCE, dep[0], dep[1], sig, [self.ptr(part_symm), self.ptr(flags_symm), self.ptr(residual),
                                                              self.ptr(residual_out), s * per, min(nvec, (s + 1) * per), 0, n_slices, s]))
        return sig, n_slices

    # stand-alone element-wise / prefetch tasks of the reference's builder (model_builder.py:336,451,494); the dense model uses the
    # fused forms (SwiGLU inside fc2's operand staging, residual add inside the all-reduce, weight prefetch inside the dispatcher)
    def _elementwise(self, ttype, ptrs, nvec, dep, extra=()):
        sig = self.counter()
        d = dep or (-1, 0)
        n_tasks = max(1, min(self.num_sms, (nvec + 2047) // 2048))
        per = (nvec + n_tasks - 1) // n_tasks
        for i in range(n_tasks):
            self._add(Task(ttype, d[0], d[1], sig, list(ptrs) + list(extra) + [i * per, min(nvec, (i + 1) * per)]))
        return sig, n_tasks

    def make_silu_mul_up(self, fc1_out, act_out, dep=None):
        """act_out[B, I] = silu(fc1_out[:, :I]) * fc1_out[:, I:]."""
        inter = fc1_out.shape[-1] // 2
        return self._elementwise(T_SILU_MUL, [self.ptr(fc1_out), self.ptr(act_out)], self.B * inter // 8, dep, extra=[inter])

    def make_add(self, lhs, rhs, output, dep=None):
        return self._elementwise(T_ADD, [self.ptr(lhs), self.ptr(rhs), self.ptr(output)], output.numel() // 8, dep)

    def make_prefetch(self, weight, dep=None):
        """DRAM -> L2 prefetch of a whole weight, spread over the CTAs (one 16 KB bulk prefetch per thread per step)."""
        sig = self.counter()
        d = dep or (-1, 0)
        kb = (weight.numel() * weight.element_size()) // 1024
        n_tasks = max(1, min(self.num_sms, (kb + 4095) // 4096))
        per = (kb + n_tasks - 1) // n_tasks
        for i in range(n_tasks):
            self._add(Task(T_PREFETCH, d[0], d[1], sig, [self.ptr(weight), i * per, max(0, min(kb, (i + 1) * per) - i * per)]))
        return sig, n_tasks

    def make_barrier_all_intra_node(self, *a, **k):
        return None   # not needed: the all-reduce tasks carry their own epoch flags

    # ---- scheduling + compile ----
    def schedule(self) -> List[List[Task]]:
        """Static per-CTA queues (``round_robin_scheduler`` / ``zig_zag_scheduler``) or one global queue in program order that CTAs
        pull from with an atomic counter at run time (``"dynamic"``)."""
        return enque_tasks(self.tasks, self.num_sms, self.schedule_policy)

    def compile(self):
        queues = self.schedule()
        self.order = [t for q in queues for t in q]
        self.task_tensor, self.queue_off = work_queue_list_to_device_tensor(queues, self.device)
        self.ptr_tensor = torch.tensor([t.data_ptr() for t in self.ptrs], dtype=torch.int64, device=self.device)
        self.sb = torch.zeros(max(self.n_counters, 1), dtype=torch.int32, device=self.device)
        self.epoch = torch.zeros(4, dtype=torch.int32, device=self.device)
        self.compiled = True
        return self

    def get_sm_activity(self) -> Dict[str, int]:
        return dict(self.metrics, tasks=len(self.tasks), counters=self.n_counters, ctas=self.num_sms)

    # ---- run ----
    def run(self, stream=None):
        assert self.compiled
        if not self.is_cuda:
            return self._run_host(getattr(self, "host_shuffle_seed", None))
        a = _MegaArgs()
        a.symm = symm_args()
        a.tasks, a.queue_off, a.ptrs = self.task_tensor.data_ptr(), self.queue_off.data_ptr(), self.ptr_tensor.data_ptr()
        a.sb, a.epoch = self.sb.data_ptr(), self.epoch.data_ptr()
        a.B, a.grid, a.smem_bytes = self.B, self.num_sms, self.max_smem
        # bits 1 / 2 select the kernel instantiation that also interprets the prefill / paged-KV task types (the plain decode
        # instantiation contains only the task bodies that have run on hardware)
        has_paged = any(t.type in (T_QKROPE_PAGED, T_ATTN_PAGED) for t in self.tasks)
        a.dynamic = int(self.schedule_policy == "dynamic") | (2 if getattr(self, "has_prefill", False) else 0) | (4 if has_paged else 0)
        a.num_tasks = len(self.tasks)
        _C.check(_C.cuda_lib().td_mega_launch(C.byref(a), C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)), "td_mega_launch")

    # emulation: interpret the task list with torch ops on the same buffers -- in program order, or (``shuffle_seed``) in a random order that
    # respects nothing but the scoreboard, which is what the GPU guarantees: a missing dependency shows up as a wrong result here
    def _run_host(self, shuffle_seed: Optional[int] = None):
        self.host_epoch = getattr(self, "host_epoch", 0) + 1
        if shuffle_seed is None:
            for t in self.tasks:
                self._host_task(t)
            return
        import random
        rng = random.Random(shuffle_seed)
        sb = [0] * max(self.n_counters, 1)
        pending = list(range(len(self.tasks)))
        while pending:
            ready = [i for i in pending if self.tasks[i].dep_idx < 0 or sb[self.tasks[i].dep_idx] >= self.tasks[i].dep_count]
            if not ready:
                raise RuntimeError(f"megakernel task graph deadlocks: {len(pending)} tasks wait on counters that are never reached")
            i = ready[-1] if shuffle_seed < 0 else rng.choice(ready)       # seed < 0: always the LATEST ready task (adversarial order)
            pending.remove(i)
            self._host_task(self.tasks[i])
            if self.tasks[i].sig_idx >= 0:
                sb[self.tasks[i].sig_idx] += 1

    def _host_task(self, t: Task):
        from ..ops.elementwise import rope_reference
        import ctypes
        heap = U.get_heap()
        lib = _C.host_lib()
        P = self.ptrs
        B = self.B
        a = t.args
        if t.type == T_RMSNORM:
            x = P[a[0]].view(B, -1).float()
            if a[1] >= 0:
                x = x + P[a[1]].view(B, -1).float()
                if a[4] >= 0:
                    P[a[4]].view(B, -1).copy_(x.to(P[a[4]].dtype))
            eps = struct.unpack("f", struct.pack("i", a[6]))[0]
            y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * P[a[2]].float()
            P[a[3]].view(B, -1).copy_(y.to(P[a[3]].dtype))
        elif t.type == T_LINEAR:
            K, n0, nc, act = a[3], a[5], a[6], a[7]
            x = P[a[0]].view(B, -1).float()
            if act == 1:
                x = torch.nn.functional.silu(x[:, :K]) * x[:, K:2 * K]
            elif act == 2:
                eps = struct.unpack("f", struct.pack("i", a[10]))[0]
                x = x[:, :K]
                x = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * P[a[9]].float()).to(P[a[0]].dtype).float()
            else:
                x = x[:, :K]
            P[a[2]].view(B, -1)[:, n0:n0 + nc] = (x @ P[a[1]][n0:n0 + nc].float().t()).to(P[a[2]].dtype)
        elif t.type in (T_QKROPE, T_QKROPE_PAGED):
            Hq, Hkv = a[7], a[8]
            eps = struct.unpack("f", struct.pack("i", a[10]))[0]
            theta = struct.unpack("f", struct.pack("i", a[11]))[0]
            qkv = P[a[0]].view(B, Hq + 2 * Hkv, -1)
            q, k, v = qkv[:, :Hq], qkv[:, Hq:Hq + Hkv], qkv[:, Hq + Hkv:]
            nrm = lambda z, w: ((z.float() * torch.rsqrt(z.float().pow(2).mean(-1, keepdim=True) + eps)) * w.float()).to(z.dtype)
            if a[4] >= 0:
                q, k = nrm(q, P[a[4]]), nrm(k, P[a[5]])
            pos = P[a[6]].view(-1)[:B]
            q, k = rope_reference(q, pos, theta), rope_reference(k, pos, theta)
            P[a[1]].view(B, Hq, -1).copy_(q)
            for b in range(B):
                if t.type == T_QKROPE_PAGED:
                    ps_, bt = a[9] & 0xFFFF, P[a[3] + 1]
                    page, slot = int(bt[b, int(pos[b]) // ps_]), int(pos[b]) % ps_
                    P[a[2]][page, slot] = k[b]
                    P[a[3]][page, slot] = v[b]
                else:
                    P[a[2]][b, int(pos[b])] = k[b]
                    P[a[3]][b, int(pos[b])] = v[b]
        elif t.type in (T_ATTN, T_ATTN_PAGED):
            b, kvh, Hq, Hkv = a[5], a[6], a[7], a[8]
            G = Hq // Hkv
            scale = struct.unpack("f", struct.pack("i", a[10]))[0]
            L = int(P[a[3]].view(-1)[b]) + 1
            q = P[a[0]].view(B, Hq, -1)[b, kvh * G:(kvh + 1) * G].float()
            split, ns = (a[11] & 0xFFFF, max(1, a[11] >> 16)) if len(a) > 11 else (0, 1)
            per = (L + ns - 1) // ns
            j0, j1 = split * per, min(L, split * per + per)
            if t.type == T_ATTN_PAGED:
                ps_, bt = a[9] & 0xFFFF, P[a[2] + 1]
                jj = torch.arange(j0, j1)
                pages, slots = bt[b, jj // ps_].long(), jj % ps_
                k, v = P[a[1]][pages, slots, kvh].float(), P[a[2]][pages, slots, kvh].float()
            else:
                k, v = P[a[1]][b, j0:j1, kvh].float(), P[a[2]][b, j0:j1, kvh].float()
            if ns == 1:
                pr = torch.softmax(q @ k.t() * scale, -1)
                P[a[4]].view(B, Hq, -1)[b, kvh * G:(kvh + 1) * G] = (pr @ v).to(P[a[4]].dtype)
            else:
                part = P[a[4]].view(-1, 130)
                sc = q @ k.t() * scale if j1 > j0 else torch.empty(G, 0)
                mm = sc.max(-1).values if j1 > j0 else torch.full((G,), float("-inf"))
                e = torch.exp(sc - mm[:, None]) if j1 > j0 else sc
                base = ((b * Hkv + kvh) * ns + split) * 8
                part[base:base + G, 0] = mm
                part[base:base + G, 1] = e.sum(-1) if j1 > j0 else 0.0
                part[base:base + G, 2:] = (e @ v) if j1 > j0 else 0.0
        elif t.type == T_QKROPE_SPLIT:
            Hq, Hkv, S = a[7], a[8], a[9]
            eps = struct.unpack("f", struct.pack("i", a[10]))[0]
            theta = struct.unpack("f", struct.pack("i", a[11]))[0]
            qkv = P[a[0]].view(B, Hq + 2 * Hkv, -1)
            q, k, v = qkv[:, :Hq], qkv[:, Hq:Hq + Hkv], qkv[:, Hq + Hkv:]
            nrm = lambda z, w: ((z.float() * torch.rsqrt(z.float().pow(2).mean(-1, keepdim=True) + eps)) * w.float()).to(z.dtype)
            if a[4] >= 0:
                q, k = nrm(q, P[a[4]]), nrm(k, P[a[5]])
            pos = (P[a[6]].view(-1)[:B // S, None] + torch.arange(S, dtype=torch.int32)[None]).reshape(-1)
            P[a[1]].view(B, Hq, -1).copy_(rope_reference(q, pos, theta))
            P[a[2]].view(B, Hkv, -1).copy_(rope_reference(k, pos, theta))
            P[a[3]].view(B, Hkv, -1).copy_(v)
        elif t.type == T_FLASH_ATTN:
            if a[5] or (a[6] & 0xFFFFFF):
                return                                    # the (head 0, q block 0) task of a batch entry does the whole entry on the host
            b, S, Hq, Hkv = a[4], a[7], a[8] & 0xFFFF, a[8] >> 16
            causal = bool((a[6] >> 30) & 1)
            scale = struct.unpack("f", struct.pack("i", a[10]))[0]
            cap = struct.unpack("f", struct.pack("i", a[11]))[0]
            q, k, v, out = (P[a[i]][b].float() for i in range(4))                    # [S, H, 128]
            G = Hq // Hkv
            sc = torch.einsum("shd,thd->hst", q, k.repeat_interleave(G, 1)) * scale
            if cap > 0:
                sc = cap * torch.tanh(sc / cap)
            if causal:
                sc = sc.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf"))
            P[a[3]][b] = torch.einsum("hst,thd->shd", torch.softmax(sc, -1), v.repeat_interleave(G, 1)).to(P[a[3]].dtype)
        elif t.type == T_ATTN_COMBINE:
            b, kvh, Hq, Hkv, ns = a[2], a[3], a[4], a[5], a[6]
            G = Hq // Hkv
            part = P[a[0]].view(-1, 130)
            rows = torch.stack([part[((b * Hkv + kvh) * ns + s_) * 8:((b * Hkv + kvh) * ns + s_) * 8 + G] for s_ in range(ns)])   # [ns, G, 130]
            mm = rows[:, :, 0].max(0).values
            c = torch.where(torch.isinf(rows[:, :, 0]), torch.zeros_like(rows[:, :, 0]), torch.exp(rows[:, :, 0] - mm[None]))
            ll = (rows[:, :, 1] * c).sum(0)
            o = (rows[:, :, 2:] * c[:, :, None]).sum(0) / ll[:, None]
            P[a[1]].view(B, Hq, -1)[b, kvh * G:(kvh + 1) * G] = o.to(P[a[1]].dtype)
        elif t.type == T_SILU_MUL:
            inter = a[2]
            x = P[a[0]].view(B, -1).float()
            y = (torch.nn.functional.silu(x[:, :inter]) * x[:, inter:2 * inter]).to(P[a[1]].dtype)
            P[a[1]].view(-1)[a[3] * 8:a[4] * 8] = y.reshape(-1)[a[3] * 8:a[4] * 8]
        elif t.type == T_ADD:
            v0, v1 = a[3] * 8, a[4] * 8
            P[a[2]].view(-1)[v0:v1] = (P[a[0]].view(-1)[v0:v1].float() + P[a[1]].view(-1)[v0:v1].float()).to(P[a[2]].dtype)
        elif t.type == T_PREFETCH:
            pass
        elif t.type == T_ALLREDUCE:
            part, flags = P[a[0]], P[a[1]]
            v0, v1, sl = a[4] * 8, a[5] * 8, a[8]
            W, me = heap.world, heap.rank
            fl = flags.view(-1)[sl * W:(sl + 1) * W]
            for r in range(W):
                lib.tdh_notify32(ctypes.c_void_p(heap.peer_ptr(fl[me:me + 1].data_ptr(), r)), self.host_epoch, 1)
            if lib.tdh_wait32_n(ctypes.c_void_p(fl.data_ptr()), W, self.host_epoch, 1, 60_000_000):
                raise TimeoutError("megakernel all-reduce flag never arrived")
            acc = torch.zeros(v1 - v0, dtype=torch.float32)
            for r in range(W):
                acc += heap.peer_view(part, (me + r) % W).view(-1)[v0:v1].float()
            if a[2] >= 0:
                acc += P[a[2]].view(-1)[v0:v1].float()
            P[a[3]].view(-1)[v0:v1] = acc.to(P[a[3]].dtype)



def _tracked(name, reads, writes):
    """Wrap a builder op: with ``auto_deps`` and no explicit ``dep`` the dependency is inferred from the named tensor arguments."""
    import functools
    import inspect
    fn = getattr(ModelBuilder, name)
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def op(self, *args, **kwargs):
        ba = sig.bind_partial(self, *args, **kwargs)
        rd = [ba.arguments.get(n) for n in reads]
        wr = [ba.arguments.get(n) for n in writes]
        if self.auto_deps and ba.arguments.get("dep") is None:
            ba.arguments["dep"] = self._infer_dep(rd, wr)
        out = fn(*ba.args, **ba.kwargs)
        self._record(out, rd, wr)
        return out
    return op


for _name, _rd, _wr in (
        ("make_rms_norm", ("x", "weight", "residual"), ("out", "residual_out")),
        ("make_linear", ("x", "weight", "norm_weight"), ("out",)),
        ("make_qk_norm_rope_update_kvcache", ("qkv", "q_norm_w", "k_norm_w", "positions", "block_table"), ("q_out", "k_cache", "v_cache")),
        ("make_flash_decode", ("q", "k_cache", "v_cache", "positions", "block_table"), ("out", "scratch")),
        ("make_flash_attn", ("q", "k", "v"), ("output",)),
        ("make_qkv_pack_qk_norm_rope_split_v", ("qkv", "kv_lens", "q_norm_w", "k_norm_w"), ("q_out", "k_out", "v_out")),
        ("make_allreduce", ("part_symm", "residual"), ("residual_out",)),
        ("make_silu_mul_up", ("fc1_out",), ("act_out",)),
        ("make_add", ("lhs", "rhs"), ("output",)),
        ("make_prefetch", ("weight",), ())):
    setattr(ModelBuilder, _name, _tracked(_name, _rd, _wr))
ModelBuilder.make_qkv_proj = ModelBuilder.make_o_proj = ModelBuilder.make_fc1 = ModelBuilder.make_fc2 = ModelBuilder.make_linear


from .dense import MegaDenseModel  # noqa: E402,F401
