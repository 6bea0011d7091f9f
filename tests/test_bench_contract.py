"""bench.py's contract on a CPU-only host: the reference arm prints ONE JSON line with the agreed keys, non-zero
ranks of a multi-rank reference launch exit silently, and our arm refuses to run without a CUDA device (there is
no CPU fallback on the product path)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gpix/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("cfg2")


def test_reference_arm_other_ranks_are_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a host without a CUDA device")
def test_our_arm_fails_loudly_without_cuda():
    r = run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
