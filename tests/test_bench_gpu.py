"""bench.py contract of our arm on the GPU box: one short run, every key the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_our_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--batch", "1024", "--cpu-budget", "2"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["metric"].startswith("SQP-RTI steps/sec") and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and d["nonzero_status"] == 0 and d["config"]["batch_per_gpu"] == 1024
    e = d["e2e"]
    assert e["ok"] and 0 < e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] == 4 * 5 and d["gpu_launches_per_tick"]["count"] == 4
    assert d["gpu_launches_per_tick"]["graphs_instantiated"] >= 1          # the device loop replays tick graphs
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "steps/s" and "sample" in cb
    assert cb["like_for_like"]["gpu_forced_ipm_over_cpu"] > 0
    # the regimes the headline does not time ride along as sub-records
    sub = d["sub_records"]
    assert sub["forced_ipm"]["mean_ipm_iterations"] >= 2 and sub["forced_ipm"]["nonzero_status"] == 0
    assert len(sub["saturated_start"]["tick_ms"]) == 12 and sub["saturated_start"]["nonzero_status"] == 0
    assert [p["horizon"] for p in sub["config5_horizon_sweep"]["points"]] == [10, 20, 40, 80]
    assert sub["config3_dob"]["value"] > 0 and sub["config3_dob"]["nonzero_status"] == 0
    assert d["e2e_explicit_yref"]["ok"] and d["e2e_explicit_yref"]["h2d_bytes_per_step"] > d["e2e"]["h2d_bytes_per_step"]


def test_both_arms_print_the_same_config():
    a = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--quick", "--no-cpu"],
                       capture_output=True, text=True, timeout=600)
    b = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-budget", "1"],
                       capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, a.stderr[-2000:] + b.stderr[-2000:]
    da, db = json.loads(a.stdout.strip().splitlines()[-1]), json.loads(b.stdout.strip().splitlines()[-1])
    assert da["config"] == db["config"] and da["metric"] == db["metric"] and da["unit"] == db["unit"]
