"""Multi-rank correctness on real GPUs: spawns one process per GPU (2 by default, NX_TEST_WORLD to
override) running tests/multirank_worker.py when at least two devices are visible; skipped on a
single-GPU box. What it covers: sharded reduce (sum / prod / max / min, 12 dtypes, NaN-sticky),
sharded argmax / argmin (ties across ranks, NaNs, the fused peer-memory finish and the all-gather
path past its limit), gathered batch-leading matmul, gradient averaging (blocking and bucketed on
the communication stream) and a captured sharded step replayed over refreshed inputs -- each
against the single-GPU answer of the same backend (reference: packages/kaun/test/test_pmap_dp.ml:18
compares the multi-device trajectory with the single-device one the same way)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world(world, extra_env=None):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-6000:]
    assert "multirank ok" in r.stdout, r.stdout[-3000:]
    return r.stdout


def _world():
    import torch
    n = torch.cuda.device_count()
    want = int(os.environ.get("NX_TEST_WORLD", "2"))
    if n < 2:
        pytest.skip(f"{n} GPU visible: the multi-rank tests need at least 2")
    return min(want, n)


@pytest.mark.gpu
def test_multirank_peer_memory():
    out = _run_world(_world())
    assert "p2p=True" in out or "p2p=False" in out


@pytest.mark.gpu
def test_multirank_nccl_only():
    """The same checks with the mailboxes disabled: every exchange goes through NCCL."""
    out = _run_world(_world(), {"NX_CUDA_P2P": "0"})
    assert "p2p=False" in out
