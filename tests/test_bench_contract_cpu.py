"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm
(`--impl reference`: the reference's own code from baseline/_ref on the host cores, else the oracle port of the same
algorithm) prints ONE JSON line with the agreed keys, under torchrun only rank 0 prints, and the GPU arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "event windows/sec voxelize+forward" and d["unit"] == "windows/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["warmup"] >= 3 and d["steps"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    want_kind = "reference" if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "learner", "learner_models.py")) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a machine without a GPU")
    r = _run(["--steps", "1"], timeout=300)
    assert r.returncode != 0      # no CPU fallback
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
