"""bench.py contract: the reference arm runs on a CPU-only box (it times the oracle port) and prints ONE JSON
line with the keys the driver reads; the GPU arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert REQUIRED <= set(out) and out["impl"] == "reference" and out["gpu_launches"] == 0
    assert out["value"] > 0 and out["unit"] == "coil-points/s" and out["higher_is_better"] is True
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0
    assert out["config"]["workload"] == "cfg1"


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = _run("--steps", "1", "--warmup", "0")
    assert res.returncode != 0  # no silent CPU fallback
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith('{"metric"')]
