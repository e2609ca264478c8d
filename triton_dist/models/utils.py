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


class MyLogger:
    """Level-keyed wrapper over the package logger (reference: models/utils.py:45-70); ``TRITON_DIST_DEBUG=1`` enables debug output."""

    def __init__(self):
        import os
        self.logger = logger
        if os.getenv("TRITON_DIST_DEBUG", "").lower() in ("true", "1", "t"):
            self.logger.setLevel(logging.DEBUG)

    def log(self, msg, level: str = "info"):
        getattr(self.logger, {"warn": "warning"}.get(level, level), self.logger.info)(msg)


def init_model_cpu(model_name: str, dtype: torch.dtype):
    """A Hugging Face causal LM on the CPU: from a LOCAL checkpoint directory (there is no network here), or -- with ``RANDOM_PARAMS=1``
    or when only a config is available -- randomly initialised from the config.  The reference downloads by name
    (models/utils.py:108-127); ``DenseLLM.init_parameters`` consumes the result the same way."""
    import os
    from transformers import AutoConfig, AutoModelForCausalLM
    with torch.no_grad():
        if os.environ.get("RANDOM_PARAMS", "0").lower() in ("1", "true", "yes") or not any(
                os.path.exists(os.path.join(model_name, f)) for f in ("model.safetensors", "pytorch_model.bin", "model.safetensors.index.json")):
            config = AutoConfig.from_pretrained(model_name)
            model = AutoModelForCausalLM.from_config(config, torch_dtype=dtype)
            return model
        return AutoModelForCausalLM.from_pretrained(model_name, torch_dtype=dtype)
