"""Protocol tests on the shared-memory emulation backend: gloo, world_size=2, no GPU (BASELINE config #1)."""
import pytest

from _launch import run_dist

CPU_ENV = {"TD_FORCE_HOST_BACKEND": "1", "CUDA_VISIBLE_DEVICES": ""}


# every distributed case at world 2; cases are grouped per launch (one torchrun + two interpreter start-ups per group instead of per case):
# the worker prints "CASE <name> OK" per case, a failure names the case in its traceback
WORLD2_GROUPS = [
    ["primitives", "allgather", "allgather_ring", "allgather_mc", "allreduce"],
    ["a2a", "ulysses_pack", "sp_pp", "sp_varlen"],
    ["ag_gemm", "gemm_rs", "gemm_ar"],
    ["gemm_a2a", "gemm_a2a_q8", "shmem"],
    ["moe", "moe_rs", "moe_staged"],
    ["ep_ll", "ep_normal", "ep_mega"],
    ["ep_fn_api", "ep_metadata", "ep_moe"],
    ["tp_e2e"],
    ["mega", "mega_paged", "engine_mega"],
    ["mega_server"],
    ["lk", "lk_shmem", "lk_ep", "lk_sp_decode", "lk_a2a"],
    ["lk_rs_ring", "lk_ar_tree", "lk_ar_push", "lk_ag_ll", "lk_ar_nvls", "allreduce_dsl", "lk_nvls_collectives"],
    ["lk_ag_gemm", "lk_gemm_ar"],
    ["lk_gemm_rs"],
]


@pytest.mark.parametrize("cases", WORLD2_GROUPS, ids=["+".join(g) for g in WORLD2_GROUPS])
def test_cpu_world2(cases):
    run_dist(cases, nproc=2, env_extra=CPU_ENV, timeout=900)


def test_cpu_world3_ring():
    run_dist(["gemm_rs", "ag_gemm"], nproc=3, env_extra=CPU_ENV)


def test_cpu_world3_fused_and_ep():
    """Non-power-of-two world: GEMM+AR, GEMM+all-to-all, both EP modes (the tiny model of the mega case needs heads % world == 0)."""
    run_dist(["gemm_ar", "gemm_a2a", "ep_ll", "ep_normal", "ep_mega", "lk", "lk_ep", "lk_rs_ring", "lk_ar_push"], nproc=3, env_extra=CPU_ENV)


def test_cpu_world4_collectives():
    run_dist(["allreduce", "allgather", "allgather_ring", "moe", "sp_pp", "shmem", "lk_ar_tree"], nproc=4, env_extra=CPU_ENV)


def test_cpu_chaos():
    """Random 0..3 ms delays in front of every notify / wait: ranks drift apart arbitrarily between protocol steps, so parity
    double-buffering, phase counters and slot reuse of every multi-call case are exercised under skew."""
    env = dict(CPU_ENV, TD_HOST_CHAOS_US="3000")
    run_dist(["ag_gemm", "gemm_rs", "gemm_ar", "gemm_a2a", "allreduce", "ep_ll", "ep_normal", "ep_mega"], nproc=3, env_extra=env)


def test_cpu_chaos_dsl_and_shmem():
    """The DSL's fused kernels (comm CTAs + tcgen05 tiles in the pipeline model), the NVSHMEM-style mirror and the NVLS all-gather
    names under random skew at a non-power-of-two world."""
    env = dict(CPU_ENV, TD_HOST_CHAOS_US="3000", TD_HOST_TIMEOUT_US="120000000")
    run_dist(["lk_gemm_rs", "shmem", "allgather_mc", "lk", "lk_shmem", "lk_ar_nvls", "lk_gemm_ar", "lk_ep"], nproc=3, env_extra=env, timeout=900)
