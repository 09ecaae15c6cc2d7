"""bench.py contract pieces that can be checked without a GPU: the reference arm (`--impl reference`: the C port of the reference's
algorithm on the host cores) prints ONE JSON line with the keys the driver reads, same metric / unit / config as our arm; ranks
other than 0 print nothing; our arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--n", "12", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "nnz/s" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "12^3" in d["config"]["workload"] and d["value"] > 0
    # the same config dict our arm prints (the driver compares them)
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    assert d["config"] == bench.config_dict(12) and d["metric"] == bench.METRIC


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--n", "12", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--no-extras", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "GPU" in (r.stderr + r.stdout)
