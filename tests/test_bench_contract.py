"""bench.py prints ONE JSON line with the contract's keys.  CPU: the reference arm (host port) with a tiny budget.
GPU: our arm, short."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-budget-s", "0.3", "--no-reference-cuda")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "fe_path_steps_per_s" and d["unit"] == "path-steps/s" and d["value"] > 1e6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
@pytest.mark.parametrize("method,metric", [("fe", "fe_path_steps_per_s"), ("em", "em_paths_per_s")])
def test_our_arm_line(method, metric):
    d = _run("--method", method, "--steps", "3", "--warmup", "3", "--log2-paths", "20", "--cpu-budget-s", "0.3", "--no-reference-cuda")
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "kernel"} <= set(d)
    assert d["metric"] == metric and d["n_gpus"] == 1 and d["gpu_launches"] == 3 and d["value"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 16
    assert abs(d["result"]["E[X]"] - 0.1197325) < 5 * d["result"]["std_error"] + 2e-4
