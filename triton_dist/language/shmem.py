"""NVSHMEM-style API at Python level: teams, put / get / put-with-signal, signal_op / signal_wait_until, team barriers, broadcast and
fcollect -- host-initiated and stream-ordered on a GPU (the role of ``pynvshmem.putmem_signal_on_stream`` / ``putmem_on_stream`` /
``team_my_pe`` in the reference: kernels/nvidia/allgather.py:268, reduce_scatter.py:538, common_ops.py:239-241), atomics on the
shared-memory heap on the emulation backend.  The device-side counterpart with the same names and semantics is ``csrc/td/shmem.cuh``
(reference: language/extra/cuda/libnvshmem_device.py:102-990).

Everything is built from three primitives of this framework: ``U.symm_at`` (peer view of a symmetric tensor), ``dl.notify`` (release
store / add of a 32-bit flag on a peer) and ``dl.wait`` (acquire spin).  Signals are the framework's native 32-bit flags.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _C
from .. import utils as U
from . import notify, wait

CMP_EQ, CMP_NE, CMP_GT, CMP_LE, CMP_LT, CMP_GE = 0, 1, 2, 3, 4, 5          # values as in NVSHMEM
SIGNAL_SET, SIGNAL_ADD = 9, 10
_OPS = {SIGNAL_SET: "set", SIGNAL_ADD: "add", "set": "set", "add": "add"}


@dataclass(frozen=True)
class Team:
    """An arithmetic progression of PEs (``nvshmem_team_split_strided``)."""
    start: int
    stride: int
    size: int

    def index_of(self, pe: int) -> int:
        d = pe - self.start
        if d < 0 or self.stride <= 0 or d % self.stride:
            return -1
        i = d // self.stride
        return i if i < self.size else -1

    def pe(self, idx: int) -> int:
        return self.start + idx * self.stride

    @property
    def pes(self):
        return [self.pe(i) for i in range(self.size)]


def my_pe() -> int:
    return U.rank()


def n_pes() -> int:
    return U.world_size()


def team_world() -> Team:
    return Team(0, 1, U.world_size())


def team_split_strided(parent: Team, start: int, stride: int, size: int) -> Team:
    """Members ``start, start + stride, ...`` (indices inside ``parent``) as a new team (global PE numbers inside)."""
    if start < 0 or size < 1 or start + (size - 1) * stride >= parent.size:
        raise ValueError("team_split_strided: the new team does not fit into its parent")
    return Team(parent.pe(start), parent.stride * stride, size)


def team_my_pe(team: Team) -> int:
    return team.index_of(U.rank())


def team_n_pes(team: Team) -> int:
    return team.size


def team_translate_pe(src_team: Team, src_pe: int, dest_team: Team) -> int:
    if not 0 <= src_pe < src_team.size:
        return -1
    return dest_team.index_of(src_team.pe(src_pe))


def remote_ptr(t: torch.Tensor, pe: int) -> torch.Tensor:
    return U.symm_at(t, pe)


# ------------------------------------------------------------------------------------------------------------
def putmem(dst: torch.Tensor, src: torch.Tensor, pe: int):
    """``src`` -> the first ``src.numel()`` elements of ``dst`` on ``pe`` (stream-ordered on a GPU).  nbi == blocking here."""
    U.symm_at(dst, pe).view(-1)[:src.numel()].copy_(src.reshape(-1))


def getmem(dst: torch.Tensor, src: torch.Tensor, pe: int):
    dst.view(-1).copy_(U.symm_at(src, pe).view(-1)[:dst.numel()])


putmem_nbi = putmem_on_stream = putmem_rma = putmem
getmem_nbi = getmem


def signal_op(sig: torch.Tensor, value: int, op, pe: int):
    notify(sig, pe, signal=value, sig_op=_OPS[op])


def quiet():
    """All my earlier puts are complete (host backend: a full fence; GPU: stream order + the release in signal_op give the guarantee)."""
    if not torch.cuda.is_available() or U.backend() != "cuda":
        _C.host_lib().tdh_fence()


fence = quiet


def putmem_signal(dst: torch.Tensor, src: torch.Tensor, sig: torch.Tensor, sig_val: int, sig_op, pe: int):
    """Data, then the flag with release semantics: whoever acquires the flag sees the data."""
    putmem(dst, src, pe)
    signal_op(sig, sig_val, sig_op, pe)


putmem_signal_nbi = putmem_signal_on_stream = putmem_signal_rma = putmem_signal


def signal_wait_until(sig: torch.Tensor, cmp: int, value: int, timeout_s: float = 60.0) -> int:
    """Block the stream (GPU: EQ / GE only, that is what the wait kernel implements) or the thread until ``sig[0] cmp value``."""
    if cmp == CMP_EQ:
        return wait(sig, 1, wait_value=value)
    if cmp == CMP_GE:
        return wait(sig, 1, wait_value=value, geq=True)
    if sig.is_cuda:
        raise NotImplementedError("stream-ordered signal_wait_until supports CMP_EQ / CMP_GE")
    test = {CMP_NE: lambda v: v != value, CMP_GT: lambda v: v > value, CMP_LE: lambda v: v <= value, CMP_LT: lambda v: v < value}[cmp]
    lib = _C.host_lib()
    t0 = time.time()
    while True:
        v = int(lib.tdh_ld_acquire32(sig.data_ptr()))
        if test(v):
            return v
        if time.time() - t0 > timeout_s:
            raise TimeoutError("signal_wait_until timed out (hang detected)")
        time.sleep(0)


# ------------------------------------------------------------------------------------------------------------
class Sync:
    """Barrier state: ``slots`` = int32 [2, world] on the symmetric heap; the epoch is tracked by this (host-side) object, which is
    valid because every collective here is host-initiated."""

    def __init__(self):
        self.slots = U.nvshmem_create_tensor((2 * U.world_size(),), torch.int32)
        self.slots.zero_()
        self.epoch = 0
        U.barrier_all_on_stream()

    def finalize(self):
        U.nvshmem_free_tensor_sync(self.slots)


def team_sync(team: Team, sync: Sync):
    """Arrival exchange among the members (flag-flip on the epoch parity; everything this rank put before is visible to a member
    once that member leaves the barrier).  Non-members return immediately."""
    sync.epoch += 1
    if team.index_of(U.rank()) < 0:
        return
    W, me, e = U.world_size(), U.rank(), sync.epoch
    arr = sync.slots[(e & 1) * W:(e & 1) * W + W]
    for pe in team.pes:
        notify(arr[me:me + 1], pe, signal=e, sig_op="set")
    for pe in team.pes:
        wait(arr[pe:pe + 1], 1, wait_value=e, geq=True)


barrier = team_sync_block = team_sync_warp = team_sync


def sync_all(sync: Sync):
    team_sync(team_world(), sync)


barrier_all = sync_all


def broadcast(team: Team, sync: Sync, dst: torch.Tensor, src: torch.Tensor, root: int):
    """The root member's ``src`` arrives in the symmetric ``dst`` of every member (root included)."""
    if team_my_pe(team) == root:
        for pe in team.pes:
            putmem(dst, src, pe)
    team_sync(team, sync)


broadcastmem = broadcast_block = broadcast_warp = broadcast


def fcollect(team: Team, sync: Sync, dst: torch.Tensor, src: torch.Tensor):
    """All-gather of equal contributions: member i's ``src`` lands at ``dst[i * n : (i + 1) * n]`` on every member."""
    me = team_my_pe(team)
    if me >= 0:
        n = src.numel()
        for q in range(team.size):
            pe = team.pe((me + q) % team.size)
            U.symm_at(dst, pe).view(-1)[me * n:(me + 1) * n].copy_(src.reshape(-1))
    team_sync(team, sync)


fcollect_block = fcollect_warp = fcollect
