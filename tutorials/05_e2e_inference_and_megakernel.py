"""Tutorial 05 -- tensor-parallel inference demo and the megakernel (reference: test_e2e_inference.py, docs/e2e.md,
mega_triton_kernel).  Random-init weights of a public architecture (no network here); pass --model Qwen/Qwen3-8B on GPUs."""
import argparse
import torch
import triton_dist.utils as U
from triton_dist.models import Engine, ModelConfig

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="tiny-dense"); ap.add_argument("--backend", default="triton_dist_gemm_ar")
ap.add_argument("--bsz", type=int, default=2); ap.add_argument("--gen_len", type=int, default=8)
args = ap.parse_args()
U.initialize_distributed()
W, me = U.world_size(), U.rank()
dt = torch.bfloat16 if U.current_device().type == "cuda" else torch.float32
eng = Engine(ModelConfig(model_name=args.model, max_length=256, dtype=dt, rank=me, world_size=W), temperature=0.0, verbose=True)
ids = torch.randint(0, 1000, (args.bsz * W, 8), generator=torch.Generator().manual_seed(0))
ref = eng.serve(ids, args.gen_len, backend="torch", use_cuda_graph=False)
out = eng.serve(ids, args.gen_len, backend=args.backend)
U.dist_print(f"backend {args.backend}: {(out == ref).float().mean().item() * 100:.0f}% of greedy tokens equal to the NCCL backend; "
             f"{eng.last_decode_ms:.3f} ms/step", allowed_ranks=[0])
U.finalize_distributed()
