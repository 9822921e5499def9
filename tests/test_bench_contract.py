"""bench.py's JSON contract: the CPU reference arm here, the CUDA arm on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run("--impl", "reference", "--steps", "4", "--warmup", "1", "--sessions", "4096")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "session_phase_steps_per_sec" and d["unit"] == "session-phase-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 4 and d["warmup"] == 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


@pytest.mark.gpu
def test_cuda_arm_prints_one_contract_line():
    d = _run("--steps", "8", "--warmup", "3", "--sessions", "65536", "--cpu-seconds", "0.5", "--e2e-calls", "2")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "fused"} <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 8 and d["gpu_launches"] >= 64          # 8 ring passes x 8 batches
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 65536 * 32 and e["d2h_bytes_per_step"] >= 65536 * 32 and e["wire"]["record_bytes"] == 32
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
