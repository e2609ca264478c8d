"""A small text-generation server on the megakernel decode step, plus its client.

Parity: the reference ships ``mega_triton_kernel/test/models/model_server.py`` (rank 0 accepts prompts on a TCP socket, broadcasts the
token ids to the other ranks, every rank runs ``mega_forwrad`` token by token) and ``chat.py`` (terminal client).  This module is the same
service for this framework, redone rather than transcribed:

* framing is newline-delimited JSON (a request or a reply is one line), so prompts are not limited to one ``recv`` buffer;
* one control message per request travels from rank 0 to the other ranks with ``torch.distributed.broadcast`` (length-prefixed int64
  tensor: token ids, generation length, sampling parameters); with more than one rank the token sampled by rank 0 is broadcast every
  step, so ranks cannot drift apart on a sampling tie caused by the different summation orders of a one-shot all-reduce;
* the prompt is prefilled either by the per-op model in one forward pass (``prefill="per_op"``, dense KV cache) or token by token through
  the megakernel (``prefill="stepwise"``: the reference's behaviour, and the only choice with a paged KV cache);
* requests: ``{"prompt": str}`` or ``{"prompt_ids": [int]}`` with optional ``max_new_tokens``, ``temperature``, ``top_p``, ``seed``;
  ``{"cmd": "stats"}``; ``{"cmd": "shutdown"}``.  Replies carry ``status``, ``response``, ``token_ids``, token counts, the device-timed
  ``processing_time`` (seconds) and ``tokens_per_s``.

Run (one process per GPU):  ``torchrun --nproc-per-node 8 -m triton_dist.mega_kernel.server --model Qwen/Qwen3-8B --port 9999``
Client:                     ``python -m triton_dist.mega_kernel.server --chat --port 9999``
"""
from __future__ import annotations

import argparse
import json
import socket
import struct
import time
from typing import List, Optional

import torch

from .. import utils as U

_HDR = 8                       # control tensor: [kind, n_prompt, max_new, seed, temperature bits, top_p bits, reserved, reserved] + ids
_KIND_GENERATE, _KIND_SHUTDOWN = 1, 2


def _fbits(x: float) -> int:
    return struct.unpack("i", struct.pack("f", float(x)))[0]


def _bits_f(i: int) -> float:
    return struct.unpack("f", struct.pack("i", int(i)))[0]


class MegaServer:
    """Owns the model, the KV cache and the megakernel task graph of ONE rank; ``serve_forever`` runs the request loop."""

    def __init__(self, model_name: str = "Qwen/Qwen3-8B", max_length: int = 4096, dtype: torch.dtype = torch.bfloat16, port: int = 9999,
                 host: str = "127.0.0.1", max_prompt: int = 1024, temperature: float = 0.6, top_p: float = 0.95, prefill: str = "per_op",
                 page_size: int = 0, eos_token_id: Optional[int] = None, model=None):
        from ..models import AutoLLM, AutoTokenizer, KV_Cache, ModelConfig, PagedKVCache
        from . import MegaDenseModel
        self.rank, self.world = U.rank(), U.world_size()
        self.host, self.port, self.max_prompt = host, port, max_prompt
        self.temperature, self.top_p = temperature, top_p
        cfg = ModelConfig(model_name=model_name, max_length=max_length, dtype=dtype, rank=self.rank, world_size=self.world)
        self.model = model if model is not None else AutoLLM.from_pretrained(cfg, U.get_triton_dist_world())
        self.tokenizer = AutoTokenizer.from_pretrained(cfg)
        m = self.model
        dev = m.device
        if page_size > 0:
            self.kv = PagedKVCache(PAGE_SIZE=page_size, num_layers=m.num_layers, batch_size=1, max_length=max_length,
                                   num_kv_heads=max(1, m.num_key_value_heads // self.world), head_dim=m.head_dim, dtype=dtype, device=dev)
            prefill = "stepwise"
        else:
            self.kv = KV_Cache(m.num_layers, 1, max_length, m.num_key_value_heads, m.head_dim, dtype, self.world, dev)
        assert prefill in ("per_op", "stepwise")
        self.prefill = prefill
        self.mega = MegaDenseModel(m, 1, self.kv)
        self.eos = eos_token_id if eos_token_id is not None else getattr(self.tokenizer, "eos_token_id", None)
        self.device = dev
        self.stats = {"requests": 0, "prompt_tokens": 0, "generated_tokens": 0, "busy_s": 0.0}
        self._sock = None

    # ---- generation (runs on every rank with identical arguments) ----
    def _reset_cache(self):
        if hasattr(self.kv, "clear"):
            self.kv.clear()
        else:
            self.kv.kv_lens.zero_()

    def _sync_token(self, tok: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            t = tok.to(torch.int64).contiguous()
            torch.distributed.broadcast(t, src=0)
            return t
        return tok

    @torch.inference_mode()
    def generate(self, prompt_ids: List[int], max_new_tokens: int = 256, temperature: Optional[float] = None, top_p: Optional[float] = None,
                 seed: int = 0):
        """-> (generated token ids, seconds).  Timed on the device when there is one (CUDA events), by the host clock on the emulation."""
        from ..models.utils import sample_token
        temperature = self.temperature if temperature is None else temperature
        top_p = self.top_p if top_p is None else top_p
        n = len(prompt_ids)
        room = self.kv.max_length - n
        if n == 0 or room <= 0:
            raise ValueError(f"prompt of {n} tokens does not fit a KV cache of {self.kv.max_length}")
        max_new_tokens = max(1, min(max_new_tokens, room))
        ids = torch.tensor(prompt_ids, dtype=torch.int64, device=self.device).view(1, -1)
        torch.manual_seed(seed)
        self._reset_cache()
        cuda = self.device.type == "cuda" if isinstance(self.device, torch.device) else str(self.device).startswith("cuda")
        if cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        t0 = time.perf_counter()
        if self.prefill == "per_op":
            pos = torch.arange(n, dtype=torch.int64, device=self.device).view(1, -1)
            logits = self.model.inference(ids, pos, self.kv)
            self.kv.inc_offset(n)
            logits = logits.view(1, -1, logits.shape[-1])[:, -1].float()
        else:
            for i in range(n):
                logits = self.mega.mega_forward(ids[:, i:i + 1])
                self.kv.inc_offset(1)
        out: List[int] = []
        for step in range(max_new_tokens):
            tok = self._sync_token(sample_token(logits, temperature=temperature, top_p=top_p))
            t = int(tok.view(-1)[0])
            out.append(t)
            if (self.eos is not None and t == self.eos) or step == max_new_tokens - 1:
                break
            logits = self.mega.mega_forward(tok.view(1, 1))
            self.kv.inc_offset(1)
        if cuda:
            ev1.record()
            ev1.synchronize()
            dt = ev0.elapsed_time(ev1) / 1e3
        else:
            dt = time.perf_counter() - t0
        return out, dt

    # ---- rank 0 <-> other ranks ----
    def _broadcast_control(self, kind: int = 0, prompt_ids=(), max_new: int = 0, seed: int = 0, temperature: float = 0.0, top_p: float = 1.0):
        ctl = torch.zeros(_HDR + self.max_prompt, dtype=torch.int64)
        if self.rank == 0:
            ctl[:6] = torch.tensor([kind, len(prompt_ids), max_new, seed, _fbits(temperature), _fbits(top_p)])
            ctl[_HDR:_HDR + len(prompt_ids)] = torch.tensor(list(prompt_ids), dtype=torch.int64)
        if self.world > 1:
            backend = torch.distributed.get_backend()
            buf = ctl.to(self.device) if backend == "nccl" else ctl
            torch.distributed.broadcast(buf, src=0)
            ctl = buf.cpu()
        kind, n, max_new, seed, tb, pb = (int(x) for x in ctl[:6])
        return kind, ctl[_HDR:_HDR + n].tolist(), max_new, seed, _bits_f(tb), _bits_f(pb)

    def _follow(self):
        """Ranks > 0: execute what rank 0 broadcasts until it says shutdown."""
        while True:
            kind, ids, max_new, seed, temperature, top_p = self._broadcast_control()
            if kind == _KIND_SHUTDOWN:
                return
            self.generate(ids, max_new, temperature, top_p, seed)

    # ---- rank 0: the socket side ----
    def _handle(self, req: dict) -> dict:
        if req.get("cmd") == "stats":
            return {"status": "success", **self.stats}
        if "prompt_ids" in req:
            ids = [int(i) for i in req["prompt_ids"]]
        elif req.get("prompt"):
            tk = self.tokenizer
            text = req["prompt"]
            if hasattr(tk, "apply_chat_template") and req.get("chat", True):
                try:
                    text = tk.apply_chat_template([{"role": "user", "content": text}], tokenize=False, add_generation_prompt=True)
                except Exception:
                    pass
            ids = list(tk.encode(text))
        else:
            return {"status": "error", "message": "prompt or prompt_ids is required"}
        if len(ids) > self.max_prompt:
            return {"status": "error", "message": f"prompt has {len(ids)} tokens, the server accepts {self.max_prompt}"}
        max_new = int(req.get("max_new_tokens", 256))
        temperature, top_p = float(req.get("temperature", self.temperature)), float(req.get("top_p", self.top_p))
        seed = int(req.get("seed", self.stats["requests"]))
        try:
            _, ids, max_new, seed, temperature, top_p = self._broadcast_control(_KIND_GENERATE, ids, max_new, seed, temperature, top_p)
            out, dt = self.generate(ids, max_new, temperature, top_p, seed)
        except Exception as e:                       # a bad request must not take the service down
            return {"status": "error", "message": f"{type(e).__name__}: {e}"}
        self.stats["requests"] += 1
        self.stats["prompt_tokens"] += len(ids)
        self.stats["generated_tokens"] += len(out)
        self.stats["busy_s"] += dt
        text = self.tokenizer.decode([t for t in out if t != self.eos])
        return {"status": "success", "response": text, "token_ids": out, "prompt_tokens": len(ids), "generated_tokens": len(out),
                "processing_time": dt, "tokens_per_s": (len(ids) + len(out)) / dt if dt > 0 else 0.0}

    def serve_forever(self, ready=None):
        """Rank 0 listens on (host, port) and answers one connection at a time; the other ranks follow its broadcasts.
        ``ready``: optional ``threading.Event`` set once the socket is listening (tests)."""
        if str(self.device).startswith("cuda"):
            torch.cuda.set_device(self.device)       # the serving thread may not be the one that initialised the process
        if self.rank != 0:
            return self._follow()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((self.host, self.port))
        self.port = srv.getsockname()[1]             # port 0 = let the OS choose
        srv.listen(8)
        self._sock = srv
        if ready is not None:
            ready.set()
        try:
            while True:
                conn, _ = srv.accept()
                with conn, conn.makefile("rwb") as f:
                    for line in f:
                        try:
                            req = json.loads(line.decode("utf-8"))
                        except (json.JSONDecodeError, UnicodeDecodeError):
                            reply = {"status": "error", "message": "a request is one line of JSON"}
                        else:
                            if req.get("cmd") == "shutdown":
                                f.write((json.dumps({"status": "success", "message": "bye"}) + "\n").encode("utf-8"))
                                f.flush()
                                self._broadcast_control(_KIND_SHUTDOWN)
                                return
                            reply = self._handle(req)
                        f.write((json.dumps(reply) + "\n").encode("utf-8"))
                        f.flush()
        finally:
            srv.close()

    def finalize(self):
        self.mega.finalize()


# ---- client ----
class Client:
    """One connection to a ``MegaServer``: ``Client(port=9999).ask("hello")`` -> reply dict."""

    def __init__(self, host: str = "127.0.0.1", port: int = 9999, timeout: float = 600.0):
        self.sock = socket.create_connection((host, port), timeout=timeout)
        self.f = self.sock.makefile("rwb")

    def request(self, payload: dict) -> dict:
        self.f.write((json.dumps(payload) + "\n").encode("utf-8"))
        self.f.flush()
        line = self.f.readline()
        if not line:
            raise ConnectionError("server closed the connection")
        return json.loads(line.decode("utf-8"))

    def ask(self, prompt: str, **kw) -> dict:
        return self.request({"prompt": prompt, **kw})

    def close(self):
        try:
            self.f.close()
        finally:
            self.sock.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def chat(host: str, port: int):
    """Terminal client (reference: test/models/chat.py): one prompt per line, ``/stats``, ``/shutdown``, ``/quit``."""
    with Client(host, port) as c:
        print(f"connected to {host}:{port} -- /quit to leave, /stats, /shutdown")
        while True:
            try:
                line = input("> ").strip()
            except EOFError:
                break
            if not line:
                continue
            if line == "/quit":
                break
            if line in ("/stats", "/shutdown"):
                print(c.request({"cmd": line[1:]}))
                if line == "/shutdown":
                    break
                continue
            r = c.ask(line)
            if r.get("status") == "success":
                print(r["response"])
                print(f"[{r['generated_tokens']} tokens, {r['processing_time']:.3f} s, {r['tokens_per_s']:.1f} tok/s]")
            else:
                print("error:", r.get("message"))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", default="Qwen/Qwen3-8B")
    ap.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float16", "float32"])
    ap.add_argument("--max_length", type=int, default=4096)
    ap.add_argument("--max_prompt", type=int, default=1024)
    ap.add_argument("--temperature", type=float, default=0.6)
    ap.add_argument("--top_p", type=float, default=0.95)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=9999)
    ap.add_argument("--prefill", default="per_op", choices=["per_op", "stepwise"])
    ap.add_argument("--page_size", type=int, default=0, help="> 0: paged KV cache with this page size")
    ap.add_argument("--chat", action="store_true", help="run the terminal client instead of the server")
    a = ap.parse_args(argv)
    if a.chat:
        return chat(a.host, a.port)
    U.initialize_distributed(seed=0)
    srv = MegaServer(a.model, a.max_length, getattr(torch, a.dtype), a.port, a.host, a.max_prompt, a.temperature, a.top_p, a.prefill, a.page_size)
    if srv.rank == 0:
        print(f"megakernel server: {a.model} on {srv.world} rank(s), listening on {a.host}:{a.port}", flush=True)
    try:
        srv.serve_forever()
    finally:
        srv.finalize()
        U.finalize_distributed()


if __name__ == "__main__":
    main()
