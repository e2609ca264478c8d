"""Dense KV cache (reference: /root/reference/python/triton_dist/models/kv_cache.py:29-66):
``[L, B, max_len, Hkv/W, D]`` K and V plus a device-resident ``kv_offset[B]`` so a captured decode graph sees the
advancing positions without re-capture."""
from __future__ import annotations

import torch


class KV_Cache:
    def __init__(self, num_layers: int, batch_size: int, max_length: int, kv_heads: int, head_dim: int,
                 dtype: torch.dtype = torch.bfloat16, world_size: int = 1, device=None):
        self.num_layers, self.batch_size, self.max_length = num_layers, batch_size, max_length
        self.kv_heads = max(1, kv_heads // world_size)
        self.head_dim = head_dim
        device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        shape = (num_layers, batch_size, max_length, self.kv_heads, head_dim)
        self.k_cache = torch.zeros(shape, dtype=dtype, device=device)
        self.v_cache = torch.zeros(shape, dtype=dtype, device=device)
        self.kv_offset = torch.zeros(batch_size, dtype=torch.int32, device=device)
        self._bidx = {}
        self._lens = torch.zeros(batch_size, dtype=torch.int32, device=device)
        self._host_len = 0      # host mirror of the longest sequence (eager calls only; Engine.serve bounds graph replays)

    def layer(self, idx: int):
        return self.k_cache[idx], self.v_cache[idx]

    def batch_index(self, bsz: int, q_len: int) -> torch.Tensor:
        key = (bsz, q_len)
        if key not in self._bidx:
            self._bidx[key] = torch.arange(bsz, dtype=torch.int32, device=self.kv_offset.device).repeat_interleave(q_len)
        return self._bidx[key]

    def kv_lens_after(self, q_len: int) -> torch.Tensor:
        """Valid KV length per batch entry once the current ``q_len`` tokens are appended (device tensor)."""
        torch.add(self.kv_offset, q_len, out=self._lens)
        return self._lens

    def get_kv_len(self) -> torch.Tensor:
        return self.kv_offset

    def inc_offset(self, n: int = 1):
        if self._host_len + n > self.max_length:
            raise ValueError(f"KV cache overflow: {self._host_len} + {n} tokens > max_length {self.max_length}")
        self._host_len += n
        self.kv_offset += n

    def clear(self):
        self._host_len = 0
        self.kv_offset.zero_()

    def rand_fill_kv_cache(self, offset: int):
        self.k_cache.normal_(0, 0.5)
        self.v_cache.normal_(0, 0.5)
        self._host_len = int(offset)
        self.kv_offset.fill_(offset)
