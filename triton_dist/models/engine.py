"""Serving engine: prefill, backend switch, CUDA-graph decode loop.

Reference: /root/reference/python/triton_dist/models/engine.py:37-189 (torch prefill -> switch backend + create ctx
sized max_M=bsz -> capture ONE decode step in a CUDA graph -> replay loop with sampling).  Every kernel of our
decode step is capturable: phases/epochs live in device memory, there are no host-side flag resets.
"""
from __future__ import annotations

import time
from typing import Optional

import torch
import torch.distributed as dist

from .. import utils as U
from ..ops import comm
from .config import ModelConfig
from .kv_cache import KV_Cache
from .utils import logger, sample_token


class Engine:
    def __init__(self, model_config: ModelConfig, temperature: float = 0.6, top_p: float = 0.95, verbose: bool = False,
                 group=None):
        from . import AutoLLM
        self.model_config = model_config
        self.group = group if group is not None else U.get_triton_dist_world()
        self.rank, self.world_size = model_config.rank, model_config.world_size
        self.temperature, self.top_p, self.verbose = temperature, top_p, verbose
        self.model = AutoLLM.from_pretrained(model_config, self.group)
        self.kv_cache: Optional[KV_Cache] = None
        self.backend = "torch"
        self.graph = None
        self.device = U.current_device()
        self.last_decode_ms = None

    # ---- setup ---------------------------------------------------------------------------------------------
    def _init_kv_cache(self, bsz: int):
        m = self.model
        self.kv_cache = KV_Cache(m.num_layers, bsz, m.max_length, m.num_key_value_heads, m.head_dim, m.dtype,
                                 self.world_size, self.device)

    def set_backend(self, backend: str, bsz: int, ar_method=comm.AllReduceMethod.Unknown):
        """torch | triton_dist | triton_dist_AR | triton_dist_gemm_ar | mega (the whole decode step as one persistent kernel:
        ``triton_dist.mega_kernel.MegaDenseModel`` on this engine's model and KV cache; dense models, batch <= 64)"""
        self.backend = backend
        if backend == "mega":
            from ..mega_kernel import MegaDenseModel
            from .dense import DenseLLM
            if not isinstance(self.model, DenseLLM) or type(self.model).__name__ != "DenseLLM":
                raise ValueError("backend='mega' needs a dense model (the megakernel task graph has no MoE tasks)")
            self.model.set_fwd("torch")
            if getattr(self, "mega", None) is not None:
                self.mega.finalize()
            self.mega = MegaDenseModel(self.model, bsz, self.kv_cache)
            U.barrier_all_host()
            return
        self.model.set_fwd(backend)
        if backend == "triton_dist":
            assert bsz % self.world_size == 0, "triton_dist (AG+RS) mode shards the batch over ranks"
            self.model.init_triton_dist_ctx(max_M=bsz)
        elif backend == "triton_dist_AR":
            self.model.init_triton_dist_AR_ctx(max_M=bsz, ar_method=ar_method)
        elif backend == "triton_dist_gemm_ar":
            self.model.init_triton_dist_gemm_ar_ctx(max_M=bsz)
        U.barrier_all_host()

    def _decode_step(self, ids: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        if self.backend == "mega":
            return self.mega.mega_forward(ids)          # positions: the KV cache's device-side offsets
        return self.model.inference(ids, pos, self.kv_cache)

    def _init_cuda_graph(self, static_ids: torch.Tensor, static_pos: torch.Tensor):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):                       # warm-up also JITs nothing: kernels are AOT, this sizes allocs
                self._decode_step(static_ids, static_pos)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        U.barrier_all_host()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.static_logits = self._decode_step(static_ids, static_pos)
        self.graph = g
        torch.cuda.synchronize()
        U.barrier_all_host()

    # ---- serving -------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def serve(self, input_ids: torch.Tensor, gen_len: int, backend: str = "torch", use_cuda_graph: bool = True,
              ar_method=comm.AllReduceMethod.Unknown) -> torch.Tensor:
        """input_ids: [bsz, prompt_len] (identical on all ranks) -> generated tokens [bsz, gen_len]."""
        bsz, prompt_len = input_ids.shape
        max_len = getattr(self.model, "max_length", None) or getattr(getattr(self.model, "cfg", None), "max_length", None)
        if max_len is not None and prompt_len + gen_len > max_len:
            raise ValueError(f"serve(): prompt_len + gen_len = {prompt_len + gen_len} exceeds the KV cache (max_length = {max_len})")
        W, r = self.world_size, self.rank
        input_ids = input_ids.to(self.device)
        self._init_kv_cache(bsz)
        # ---- prefill (torch mode, NCCL all-reduce; as the reference does) ----
        self.model.set_fwd("torch")
        pos = torch.arange(prompt_len, device=self.device, dtype=torch.int64)[None, :].expand(bsz, -1).contiguous()
        logits = self.model.inference(input_ids, pos, self.kv_cache)
        self.kv_cache.inc_offset(prompt_len)
        next_tok = sample_token(logits, self.temperature, self.top_p)
        if W > 1:
            dist.broadcast(next_tok, src=dist.get_global_rank(self.group, 0), group=self.group)   # ranks must agree
        out = [next_tok]
        # ---- backend switch ----
        self.set_backend(backend, bsz, ar_method)
        sharded = backend == "triton_dist"
        b_local = bsz // W if sharded else bsz
        static_ids = torch.empty((b_local, 1), dtype=torch.int64, device=self.device)
        static_pos = torch.empty((bsz, 1), dtype=torch.int64, device=self.device)

        def load_inputs(tok):
            static_ids.copy_(tok[r * b_local:(r + 1) * b_local] if sharded else tok)
            static_pos.copy_(self.kv_cache.kv_offset.to(torch.int64)[:, None])

        load_inputs(next_tok)
        use_graph = use_cuda_graph and self.device.type == "cuda"
        if use_graph:
            self._init_cuda_graph(static_ids, static_pos)
        t0 = time.time()
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        # optional torch profiler over the first ``profile_steps`` decode steps (reference: engine.py:149-182, `enable_profile`, 64 steps):
        # one chrome trace per rank under ``profile_dir``
        profiler, prof_left = None, 0
        if getattr(self, "enable_profile", False):
            import os
            acts = [torch.profiler.ProfilerActivity.CPU] + ([torch.profiler.ProfilerActivity.CUDA] if self.device.type == "cuda" else [])
            profiler = torch.profiler.profile(activities=acts, record_shapes=False)
            prof_left = int(getattr(self, "profile_steps", 64))
            self._profile_dir = getattr(self, "profile_dir", "prof")
            os.makedirs(self._profile_dir, exist_ok=True)
            profiler.__enter__()

        def _stop_profile():
            nonlocal profiler
            if profiler is not None:
                if self.device.type == "cuda":
                    torch.cuda.synchronize()
                profiler.__exit__(None, None, None)
                self.last_trace = f"{self._profile_dir}/decode_{backend}_rank{r}.json"
                profiler.export_chrome_trace(self.last_trace)
                profiler = None

        for _ in range(gen_len - 1):
            if profiler is not None and prof_left == 0:
                _stop_profile()
            prof_left -= 1
            load_inputs(next_tok)
            if use_graph:
                self.graph.replay()
                logits = self.static_logits
            else:
                logits = self._decode_step(static_ids, static_pos)
            if sharded:
                full = torch.empty((bsz, logits.shape[-1]), dtype=logits.dtype, device=self.device)
                dist.all_gather_into_tensor(full, logits.contiguous(), group=self.group)
                logits = full
            next_tok = sample_token(logits, self.temperature, self.top_p)
            if W > 1:
                dist.broadcast(next_tok, src=dist.get_global_rank(self.group, 0), group=self.group)
            self.kv_cache.inc_offset(1)
            out.append(next_tok)
        _stop_profile()
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        self.last_decode_ms = (time.time() - t0) * 1e3 / max(1, gen_len - 1)
        if self.verbose and r == 0:
            logger.info(f"decode: {self.last_decode_ms:.3f} ms/step ({backend}, graph={use_graph})")
        return torch.cat(out, dim=1)

    def finalize(self):
        self.graph = None
        if getattr(self, "mega", None) is not None:
            self.mega.finalize()
            self.mega = None
        self.model.finalize()
