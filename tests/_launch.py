"""Spawn tests/dist_worker.py under torchrun (127.0.0.1 rendezvous) and surface its output on failure."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_dist(cases, nproc=2, timeout=600, env_extra=None):
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "2")
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")] + list(cases)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    if r.returncode != 0:
        raise AssertionError(f"distributed worker failed (rc={r.returncode})\n--- stdout ---\n{r.stdout[-6000:]}\n--- stderr ---\n{r.stderr[-6000:]}")
    for c in cases:
        assert f"CASE {c} OK" in r.stdout, r.stdout[-3000:]
    return r.stdout
