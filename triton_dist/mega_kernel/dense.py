"""Dense-LLM decode step on the megakernel (reference: mega_triton_kernel/models/dense.py:108-201, ``mega_forwrad``).

Builds, for every layer:  (rms_norm fused into) qkv_proj -> qk_norm_rope_update_kvcache -> split-KV flash_decode -> combine -> o_proj
-> allreduce(+residual) -> (rms_norm fused into) fc1 -> (silu*mul fused into) fc2 -> allreduce(+residual).  Embedding, final norm and lm_head stay
outside the persistent kernel (as in the reference, which copies the embedded tokens into the hidden-state buffer and
runs the builder)."""
from __future__ import annotations

import torch

from .. import utils as U
from ..models.dense import DenseLLM
from ..models.kv_cache import KV_Cache
from ..ops.elementwise import rmsnorm
from ..parallel.tp_mlp import _linear


class MegaDenseModel:
    def __init__(self, model: DenseLLM, batch: int, kv_cache: KV_Cache, num_sms=None, schedule: str = "round_robin",
                 fuse_norm: bool = True, attn_splits: int = 0):
        from . import ModelBuilder
        self.model, self.B, self.kv = model, batch, kv_cache
        a = model.arch
        dev, dt = model.device, model.dtype
        heap = U.get_heap()
        W = model.world_size
        l0 = model.layers[0]
        Hq, Hkv, D, H = l0.attn.Hq, l0.attn.Hkv, l0.attn.D, a.hidden_size
        B = batch
        L = len(model.layers)
        mb = ModelBuilder(B, num_sms, schedule)
        self.h = torch.zeros((B, H), dtype=dt, device=dev)                 # residual stream
        self.xn = torch.zeros((B, H), dtype=dt, device=dev)
        self.qkv = torch.zeros((B, (Hq + 2 * Hkv) * D), dtype=dt, device=dev)
        self.q_rot = torch.zeros((B, Hq * D), dtype=dt, device=dev)
        self.attn_out = torch.zeros((B, Hq * D), dtype=dt, device=dev)
        self.gu = torch.zeros((B, l0.mlp.gate_up_proj.shape[0]), dtype=dt, device=dev)
        self.parts = heap.tensor((2 * L, B, H), dt)                          # symmetric partial sums, one per (layer, op)
        self.n_slices = 4
        self.flags = heap.tensor((2 * L, self.n_slices, max(W, 4)), torch.int32)
        self.positions = torch.zeros(B, dtype=torch.int32, device=dev)
        # split-KV: enough attention tasks to cover the SMs (B * Hkv heads alone leave most of them idle)
        if attn_splits <= 0:
            attn_splits = max(1, min(16, mb.num_sms // max(1, B * Hkv), kv_cache.max_length // 64))
        self.attn_splits = attn_splits
        self.attn_scratch = torch.zeros((B * Hkv * attn_splits * 8 * 130,), dtype=torch.float32, device=dev) if attn_splits > 1 else None
        U.barrier_all_host()
        dep = None
        from ..models.paged_kv_cache import PagedKVCache
        self.paged = isinstance(kv_cache, PagedKVCache)
        for li, layer in enumerate(model.layers):
            mb.cur_layer = li
            if self.paged:          # one page pool for all layers, a block table per layer
                k_cache, v_cache, bt = kv_cache.key_cache, kv_cache.value_cache, kv_cache.block_tables[li].contiguous()
            else:
                (k_cache, v_cache), bt = kv_cache.layer(li), None
            at, ml = layer.attn, layer.mlp
            if fuse_norm:
                d = mb.make_qkv_proj(self.h, at.wqkv, self.qkv, dep, norm_weight=layer.input_norm_w, eps=layer.eps)
            else:
                d = mb.make_rms_norm(self.h, layer.input_norm_w, self.xn, layer.eps, dep=dep)
                d = mb.make_qkv_proj(self.xn, at.wqkv, self.qkv, d)
            d = mb.make_qk_norm_rope_update_kvcache(self.qkv, self.q_rot, k_cache, v_cache, at.q_norm_w, at.k_norm_w, self.positions,
                                                    Hq, Hkv, at.eps, at.rope_theta, d, block_table=bt)
            d = mb.make_flash_decode(self.q_rot, k_cache, v_cache, self.positions, self.attn_out, Hq, Hkv, at.sm_scale, d,
                                     n_splits=attn_splits, scratch=self.attn_scratch, block_table=bt)
            d = mb.make_o_proj(self.attn_out, at.wo, self.parts[2 * li], d)
            d = mb.make_allreduce(self.parts[2 * li], self.flags[2 * li], self.h, self.h, d, self.n_slices)
            if fuse_norm:
                d = mb.make_fc1(self.h, ml.gate_up_proj, self.gu, d, norm_weight=layer.post_norm_w, eps=layer.eps)
            else:
                d = mb.make_rms_norm(self.h, layer.post_norm_w, self.xn, layer.eps, dep=d)
                d = mb.make_fc1(self.xn, ml.gate_up_proj, self.gu, d)
            d = mb.make_fc2(self.gu, ml.down_proj, self.parts[2 * li + 1], d, act_silu_mul=True)
            dep = mb.make_allreduce(self.parts[2 * li + 1], self.flags[2 * li + 1], self.h, self.h, d, self.n_slices)
        self.builder = mb.compile()

    @torch.inference_mode()
    def mega_forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        """input_ids: [B, 1] -> fp32 logits [B, V].  Positions come from the KV cache's device-side offsets."""
        m = self.model
        self.h.copy_(torch.nn.functional.embedding(input_ids.view(-1), m.embed_tokens))
        self.positions.copy_((self.kv.kv_lens if self.paged else self.kv.kv_offset)[: self.B])
        self.builder.run()
        hn = rmsnorm(self.h, m.norm_w, m.arch.rms_norm_eps)
        return _linear(hn, m.lm_head).float()

    mega_forwrad = mega_forward     # the reference's spelling

    def finalize(self):
        heap = U.get_heap()
        heap.free_tensor(self.parts)
        heap.free_tensor(self.flags)


DenseModel = MegaDenseModel     # the reference's class name (mega_triton_kernel/models/dense.py)

