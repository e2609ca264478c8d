"""Models (reference: /root/reference/python/triton_dist/models/__init__.py:33-69)."""
from .config import ARCHS, ArchConfig, ModelConfig  # noqa: F401
from .kv_cache import KV_Cache  # noqa: F401
from .dense import DenseLLM  # noqa: F401
from .qwen_moe import Qwen3MoE  # noqa: F401
from .utils import logger, sample_token, seed_everything  # noqa: F401


class AutoLLM:
    @staticmethod
    def model_mapping():
        from .qwen_moe import Qwen3MoE
        m = {name: (Qwen3MoE if a.num_experts else DenseLLM) for name, a in ARCHS.items()}
        return m

    @staticmethod
    def from_pretrained(model_config: ModelConfig, group=None):
        cls = AutoLLM.model_mapping().get(model_config.model_name)
        if cls is None:
            import os
            if os.path.isfile(os.path.join(model_config.model_name, "config.json")):     # local HF checkpoint directory
                cls = Qwen3MoE if model_config.arch().num_experts else DenseLLM
            else:
                raise KeyError(f"unsupported model {model_config.model_name}")
        return cls(model_config, group)


class AutoTokenizer:
    """Thin wrapper over HF tokenizers when a local copy exists; a byte-level fallback otherwise (no network)."""

    @staticmethod
    def from_pretrained(model_config: ModelConfig):
        try:
            from transformers import AutoTokenizer as HF
            return HF.from_pretrained(model_config.model_name, local_files_only=True)
        except Exception:
            return _ByteTokenizer(model_config.arch().vocab_size)


class _ByteTokenizer:
    def __init__(self, vocab):
        self.vocab = vocab

    def encode(self, text):
        return [b % self.vocab for b in text.encode("utf-8")]

    def decode(self, ids):
        return bytes(int(i) % 256 for i in ids).decode("utf-8", errors="replace")


from .engine import Engine  # noqa: E402,F401
from .paged_kv_cache import PagedKVCache  # noqa: E402,F401
