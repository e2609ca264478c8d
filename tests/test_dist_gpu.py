"""Multi-GPU checks (torchrun over NCCL + our symmetric heap).  Skipped when fewer than 2 GPUs are visible."""
import pytest
import torch

from _launch import run_dist

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case", ["primitives", "allgather", "allreduce", "ag_gemm", "gemm_rs", "gemm_ar", "gemm_q8", "gemm_a2a", "moe", "moe_rs", "moe_staged", "tp_e2e", "ep_ll", "ep_normal", "ep_mega", "sp_pp", "ep_moe", "mega"])
def test_gpu_world2(case):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    run_dist([case], nproc=2, timeout=300)


@pytest.mark.parametrize("case", ["ag_gemm", "gemm_rs", "gemm_ar", "moe", "moe_rs", "ep_ll", "ep_normal", "ep_mega", "mega", "tp_e2e"])
def test_gpu_world4(case):
    if _ngpu() < 4:
        pytest.skip("needs >= 4 GPUs")
    run_dist([case], nproc=4, timeout=420)


@pytest.mark.parametrize("case", ["ag_gemm", "gemm_rs", "gemm_ar", "moe", "moe_rs", "ep_ll", "ep_normal", "ep_mega", "mega", "tp_e2e"])
def test_gpu_world8(case):
    if _ngpu() < 8:
        pytest.skip("needs >= 8 GPUs")
    run_dist([case], nproc=8, timeout=600)
