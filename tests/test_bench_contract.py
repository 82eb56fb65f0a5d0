"""bench.py contract checks that need no GPU: the reference arm (CPU port of the reference path) prints one JSON
line with the keys the driver reads, and the workload-name pattern of the batch sweep parses."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "ou_b16_t20",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "trajectory-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("trajectory-steps/s") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "ou_b16_t20"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_nonzero_rank_is_silent(monkeypatch):
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                          "ou_b16_t20", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_names():
    from viforsdes_b200.synthetic import WORKLOADS

    assert WORKLOADS["lv_b128_t800"] == ("lv", 128, 800, 0.05)
    assert WORKLOADS["ou_b4096_t100"] == ("ou", 4096, 100, 0.05)
    with pytest.raises(KeyError):
        WORKLOADS["nope_b1_t1"]
