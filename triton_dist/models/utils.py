"""Model utilities: logger, seeding, sampling (reference: /root/reference/python/triton_dist/models/utils.py)."""
from __future__ import annotations

import logging
import random
import sys

import numpy as np
import torch


class _Color(logging.Formatter):
    C = {"DEBUG": "\033[36m", "INFO": "\033[32m", "WARNING": "\033[33m", "ERROR": "\033[31m"}

    def format(self, record):
        c = self.C.get(record.levelname, "")
        return f"{c}[{record.levelname}] {record.getMessage()}\033[0m"


logger = logging.getLogger("triton_dist")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(_Color())
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)


def seed_everything(seed: int = 42):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def sample_token(logits: torch.Tensor, temperature: float = 0.6, top_p: float = 0.95, top_k: int = -1) -> torch.Tensor:
    """logits [B, V] -> next token ids [B, 1] (temperature / nucleus sampling; temperature 0 = greedy)."""
    if temperature <= 0:
        return logits.argmax(dim=-1, keepdim=True)
    probs = torch.softmax(logits.float() / temperature, dim=-1)
    sorted_p, idx = torch.sort(probs, dim=-1, descending=True)
    if top_k > 0:
        sorted_p[:, top_k:] = 0
    cum = torch.cumsum(sorted_p, dim=-1)
    sorted_p = sorted_p.masked_fill(cum - sorted_p > top_p, 0.0)
    sorted_p = sorted_p / sorted_p.sum(dim=-1, keepdim=True)
    nxt = torch.multinomial(sorted_p, 1)
    return idx.gather(-1, nxt)
