"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm prints exactly one
JSON line with the keys the driver reads, on the same workload config the GPU arm reports, and the GPU arm refuses to run
(non-zero exit, nothing on stdout) when there is no CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "images/sec fwd+loss" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # same workload description as the GPU arm (bench.workload_config is the single source of both)
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    cfg = bench.workload_config(argparse.Namespace(height=256, batch=256, scaling="weak", workload="fwdloss"), 1)
    assert d["config"] == cfg and cfg["per_gpu_batch"] == 256 and "HRNet-W32" in cfg["workload"]


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA present")
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr
