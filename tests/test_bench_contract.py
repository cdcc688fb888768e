"""bench.py's output contract, on the legs that run without a GPU: the reference arm prints exactly ONE JSON line on
stdout with the keys the driver reads, other ranks of a torchrun launch print nothing, and the own arm refuses to run
(loudly) when there is no CUDA device -- it has no CPU path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line(ref_available):
    r = run_bench(["--impl", "reference", "--config", "mono_tum", "--steps", "1", "--warmup", "3"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["warmup"] >= 3 and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = run_bench(["--impl", "reference", "--config", "mono_tum", "--gpus", "2"], env={"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_bench(["--steps", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "CUDA" in (r.stderr + r.stdout)
