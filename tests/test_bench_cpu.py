"""bench.py host logic that runs without a GPU: the reference (CPU) arm's JSON contract and the loud failure of our arm."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_contract():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-budget", "1"])
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_all_cores_under_torchrun_env():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-budget", "1"], env={"OMP_NUM_THREADS": "1"})
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_our_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is visible")
    r = _run(["--steps", "3", "--warmup", "3"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and "no CPU path" in (r.stderr + r.stdout)
