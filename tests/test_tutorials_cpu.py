"""The tutorials run end to end on the emulation backend (torchrun, gloo, 2 ranks) -- they are the first thing a user tries."""
import os
import subprocess
import sys

import pytest

from _launch import ROOT, free_port


@pytest.mark.parametrize("script,expect", [("01_notify_wait.py", "rounds OK"), ("03_ag_gemm_gemm_rs.py", "ag_gemm -> gemm_rs OK"),
                                           ("06_sequence_parallel_attention.py", "gemm + all-to-all"), ("09_mega_ep_and_fused_moe.py", "fused MoE tutorial OK"),
                                           ("10_kernel_dsl.py", "kernel DSL tutorial OK"),
                                           ("11_distributed_kernels_in_python.py", "distributed DSL kernels tutorial OK")])
def test_tutorial(script, expect):
    env = dict(os.environ, TD_FORCE_HOST_BACKEND="1", CUDA_VISIBLE_DEVICES="", TD_SYMM_HEAP_SIZE="256m", OMP_NUM_THREADS="2",
               PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tutorials", script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert expect in r.stdout + r.stderr


def test_stress_script_on_the_emulation_backend():
    """The stress test of the flagship ops (random shapes on one pair of contexts, unverified back-to-back calls, then a verified one) at a
    non-power-of-two world with random skew before every flag operation (reference: test/stress/stress_test_ag_gemm.py)."""
    env = dict(os.environ, TD_FORCE_HOST_BACKEND="1", CUDA_VISIBLE_DEVICES="", TD_SYMM_HEAP_SIZE="256m", OMP_NUM_THREADS="2", TD_HOST_CHAOS_US="1500",
               PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "triton_dist", "test", "stress", "stress_test_ag_gemm.py"),
           "--max_M", "96", "--N", "48", "--K", "64", "--iters", "2", "--verify_hang", "4", "--verify_shapes", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "stress test OK" in r.stdout + r.stderr, r.stdout[-3000:] + r.stderr[-3000:]

