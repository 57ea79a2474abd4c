"""The bench.py contract the driver depends on, as far as it can be exercised without a GPU: the
reference arm (`--impl reference`, the CPU oracle timed on the host cores) prints ONE JSON line with the
base keys plus "impl", "cpu_baseline" and a zero-copy "e2e"; ranks other than 0 print nothing and exit 0;
without a GPU the device arm refuses loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_search_line():
    r = _run(["--impl", "reference", "--workload", "search", "--steps", "1", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= d.keys() and d["impl"] == "reference" and d["vs_baseline"] is None
    assert d["unit"] == "queries/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"].startswith("exact IP search over 10M x 512")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--workload", "search", "--gpus", "2", "--steps", "1", "--warmup", "3"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_device_arm_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = _run(["--steps", "3", "--warmup", "3"])
    assert r.returncode != 0 and "no CPU path" in r.stderr and not any(l.startswith("{") for l in r.stdout.splitlines())
