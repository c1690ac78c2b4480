"""bench.py contract checks that need no GPU: the `--impl reference` arm (the reference's own CPU implementation from
oracle/_ref, or the oracle port where the reference was not built) prints ONE JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--log2n", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "elements/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--log2n", "16", "--no-cpu-baseline", "--no-primitives", "--no-sharded")
    assert r.returncode != 0, "bench.py must not produce a number without a GPU"
    assert not any(l.startswith("{") and '"value"' in l for l in r.stdout.splitlines())
