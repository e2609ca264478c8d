"""Point-to-point primitives for pipeline parallelism (reference: kernels/nvidia/p2p.py:33-150 --
``p2p_set_signal``, ``p2p_wait_signal``, ``p2p_copy_kernel`` (get), ``p2p_put_kernel`` (put + signal)).
Data moves with peer-mapped copies on the symmetric heap; signals are release/acquire flag words."""
from __future__ import annotations

import torch

from .. import language as dl
from .. import utils as U
from .comm import copy_tensor


def p2p_set_signal(signal: torch.Tensor, peer: int, value: int = 1, op: str = "set"):
    dl.notify(signal, peer, signal=value, sig_op=op)


def p2p_wait_signal(signal: torch.Tensor, value: int = 1, geq: bool = True):
    dl.wait(signal, 1, "sys", "acquire", wait_value=value, geq=geq)


def p2p_put(dst_symm: torch.Tensor, src: torch.Tensor, peer: int, signal: torch.Tensor = None, value: int = 1):
    """Write ``src`` into ``peer``'s copy of ``dst_symm`` (then raise ``signal`` there)."""
    copy_tensor(U.symm_at(dst_symm, peer)[: src.shape[0]] if dst_symm.dim() else U.symm_at(dst_symm, peer), src)
    if signal is not None:
        p2p_set_signal(signal, peer, value)


def p2p_get(dst: torch.Tensor, src_symm: torch.Tensor, peer: int):
    """Read ``peer``'s copy of ``src_symm`` into local ``dst``."""
    copy_tensor(dst, U.symm_at(src_symm, peer)[: dst.shape[0]] if src_symm.dim() else U.symm_at(src_symm, peer))


p2p_copy_remote_to_local = p2p_get
# the reference names its device kernels (p2p.py:66,89,119); here a transfer is one call of the vectorised copy kernel on peer-mapped memory
p2p_copy_kernel = p2p_get
p2p_copy_remote_to_local_kernel = p2p_get
p2p_put_kernel = p2p_put

