"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): runs without a GPU, prints one JSON line
with the keys the contract names, uses every host core it may run on, and (ranks > 0 of a multi-rank launch) exits
without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *args],
                          capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))


def test_reference_arm_line_has_the_contract_keys():
    res = _run({"OMP_NUM_THREADS": "1"})  # what torch.distributed.run exports to its workers: must not decide the thread count
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "atom-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["value"] == d["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))  # every core, although OMP_NUM_THREADS=1 was exported
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_ranks_above_zero_exit_without_work():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_refuses_to_run_without_a_gpu():
    """The product arm has no CPU path: without a device it must stop with a message, not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=str(ROOT))
    assert res.returncode != 0
    assert "needs a GPU" in (res.stderr + res.stdout)
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]
