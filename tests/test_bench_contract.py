"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU oracle timed on the host
cores) prints exactly ONE JSON line with the agreed keys, ranks other than 0 stay silent, and the GPU arm refuses to run
without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sky-pixels/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    r = run(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0", "--gpus", "2"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("checks the no-GPU behaviour")
    r = run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
