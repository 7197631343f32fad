"""bench.py contract (CPU side): the reference arm prints ONE JSON line with the agreed keys, and the CUDA arm
refuses to run without a GPU instead of falling back to anything (the product path has no CPU route)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mel-frames/sec" and d["unit"] == "mel-frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] == d["value"] and c["sample"]
    # a step is a bounded sample (2 of the 31 Euler intervals), extrapolated to the whole utterance
    assert abs(d["value"] - 937 / (d["ms_per_step"] / 1e3 * 15.5)) / d["value"] < 1e-6


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_cuda_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert r.returncode != 0
    assert r.stdout.strip() == ""                       # no JSON line: nothing was measured
    assert "no CUDA device" in r.stderr
