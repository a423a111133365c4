"""Multi-GPU invariance through the real launch path: these tests spawn `torch.distributed.run` on the GPUs of the box
(skipped when it has fewer than two) and run tests/mgpu_check.py, which compares the N-rank trajectory with the
single-GPU one (bitwise for the replicated-coordinates and the brick-decomposed engines, 1e-10 for the NCCL halo engine)
and with the CPU oracle.  The one-rank cases of the brick engine run everywhere."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _gpus() -> int:
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _spawn(world: int, args, timeout=600):
    env = dict(os.environ)
    env.setdefault("PANTEA_DIST_TIMEOUT_S", "120")
    import socket
    with socket.socket() as sock:  # a port that is free right now (consecutive launches do not wait out TIME_WAIT)
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "mgpu_check.py"), *[str(a) for a in args]]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@pytest.mark.parametrize("n_atoms,steps", [(12000, 6), (24000, 30)])
def test_brick_engine_one_rank_is_the_single_gpu_loop(n_atoms, steps):
    """world = 1: mailbox, roles and graph replay on one GPU; bitwise equal to pantea_md_run, forces vs the oracle."""
    res = subprocess.run([sys.executable, str(ROOT / "tests" / "mgpu_check.py"), str(n_atoms), str(steps), "brick", "oracle"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "bitwise_identical=True" in res.stdout and "one_owner=True" in res.stdout


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_brick_engine_n_ranks_bitwise_and_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _spawn(world, [24000, 20, "brick", "oracle"])
    assert "bitwise_identical=True" in out and "one_owner=True" in out and "neighbour_counts_equal=True" in out


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_replicated_engine_two_ranks_bitwise():
    out = _spawn(2, [24000, 5])
    assert "bitwise_identical=True" in out


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_halo_engine_two_ranks():
    out = _spawn(2, [24000, 6, "halo"])
    assert "halo world=2" in out
