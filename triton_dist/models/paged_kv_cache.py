"""Paged KV cache (reference: python/triton_dist/mega_triton_kernel/models/paged_kv_cache.py:28-58 -- pages of ``PAGE_SIZE`` tokens, a
random page permutation per (layer, sequence), ``get_layer_kv_cache`` -> (key pages, value pages, block table, kv lengths)).

The flash-decode kernel of this framework reads through block tables (``ops.flash_decode.gqa_fwd_batch_decode(..., block_table=)``,
csrc/attention.cu), so a layer's cache plugs straight into it; ``append`` writes new tokens through the table (what the fused
qk-norm + rope + KV-store kernel does for the dense cache)."""
from __future__ import annotations

import torch


class PagedKVCache:
    def __init__(self, PAGE_SIZE: int = 16, num_layers: int = 32, batch_size: int = 1, max_length: int = 1024, num_kv_heads: int = 8,
                 head_dim: int = 128, dtype=torch.bfloat16, device=None, seed: int = 0) -> None:
        dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        self.page_size = PAGE_SIZE
        self.max_num_blocks_per_seq = (max_length + PAGE_SIZE - 1) // PAGE_SIZE
        self.num_layers, self.batch_size, self.max_length = num_layers, batch_size, max_length
        self.num_kv_heads, self.head_dim, self.dtype = num_kv_heads, head_dim, dtype
        n_blocks = self.max_num_blocks_per_seq * batch_size * num_layers
        self.key_cache = torch.zeros(n_blocks, PAGE_SIZE, num_kv_heads, head_dim, dtype=dtype, device=dev)
        self.value_cache = torch.zeros(n_blocks, PAGE_SIZE, num_kv_heads, head_dim, dtype=dtype, device=dev)
        g = torch.Generator().manual_seed(seed)          # pages are deliberately scattered: adjacent tokens are not adjacent in memory
        self.block_tables = torch.randperm(n_blocks, generator=g, dtype=torch.int64).to(torch.int32).to(dev).reshape(
            num_layers, batch_size, self.max_num_blocks_per_seq)
        self.kv_lens = torch.zeros(batch_size, dtype=torch.int32, device=dev)

    def inc_offset(self, seq_len: int):
        if int(self.kv_lens.max()) + seq_len > self.max_length:
            raise ValueError("PagedKVCache: sequence longer than max_length")
        self.kv_lens += seq_len

    def get_layer_kv_cache(self, layer_idx: int):
        return self.key_cache, self.value_cache, self.block_tables[layer_idx], self.kv_lens

    def append(self, layer_idx: int, k_new: torch.Tensor, v_new: torch.Tensor):
        """k_new / v_new: [B, S, Hkv, D] -> positions ``kv_lens[b] .. + S`` of every sequence (call ``inc_offset(S)`` after the last layer)."""
        B, S = k_new.shape[:2]
        pos = self.kv_lens.long()[:, None] + torch.arange(S, device=k_new.device)[None]            # [B, S]
        if int(pos.max()) >= self.max_length:
            raise ValueError("PagedKVCache: sequence longer than max_length")
        page = torch.gather(self.block_tables[layer_idx].long(), 1, pos // self.page_size)            # [B, S]
        slot = pos % self.page_size
        self.key_cache[page, slot] = k_new.to(self.dtype)
        self.value_cache[page, slot] = v_new.to(self.dtype)

    def gather_dense(self, layer_idx: int):
        """[B, max_len, Hkv, D] dense views of one layer (tests / fallbacks)."""
        bt = self.block_tables[layer_idx].long()
        k = self.key_cache[bt].reshape(self.batch_size, -1, self.num_kv_heads, self.head_dim)[:, :self.max_length]
        v = self.value_cache[bt].reshape(self.batch_size, -1, self.num_kv_heads, self.head_dim)[:, :self.max_length]
        return k, v
