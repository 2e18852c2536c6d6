"""bench.py's output contract, checked on the leg that needs no GPU: `--impl reference` times the oracle
port of the reference's CPU path and must print exactly ONE JSON line on stdout with the agreed keys;
the B200 arm must refuse to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"}


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=280, env=dict(os.environ, **(env or {})))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1", "--cpu-seconds", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    j = json.loads(lines[0])
    assert KEYS <= set(j), KEYS - set(j)
    assert j["impl"] == "reference" and j["unit"] == "pairs/s" and j["higher_is_better"] is True
    assert j["vs_baseline"] is None and j["steps"] == 2 and j["value"] > 0
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "candidates" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.timeout(300)
def test_reference_arm_non_zero_ranks_exit_quietly():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
@pytest.mark.timeout(120)
def test_b200_arm_has_no_cpu_fallback():
    r = _run("--workload", "tiny", "--steps", "1")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr
