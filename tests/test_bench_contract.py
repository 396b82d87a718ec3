"""bench.py's JSON contract on the parts that run without a GPU: the reference arm (oracle port on the host
cores) and the refusal of our arm to run without a device."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def run(*args, env=None):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600, env=env)


def test_reference_arm_prints_one_contract_line():
    p = run("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1", "--ref-seconds", "3")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s (all bounces)" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("diffuse_single_sphere") or "512" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "spp" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_zero_only():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = run("--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1", "--warmup", "0", env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    p = run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
