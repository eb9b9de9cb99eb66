"""CPU tests of bench.py's contract: the reference arm (reference CPU implementation on the host cores) prints one JSON
line with the agreed keys for any --steps, and our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.pop("RANK", None); e.pop("WORLD_SIZE", None)
    if env:
        e.update(env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


@pytest.mark.parametrize("steps", [2, 15])
def test_reference_arm_line(steps):
    r = run_bench("--impl", "reference", "--grid", "32", "--steps", str(steps), "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["higher_is_better"] is True
    assert d["steps"] == steps and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "poisson3d_32" and d["config"]["timed_iterations"] >= 10
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a rate, not a timer artefact: 32^3 BiCG iterations take 0.1 ms .. 1 s each on any host
    assert 1.0 < d["value"] < 1e5 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6


def test_reference_arm_other_ranks_are_silent():
    r = run_bench("--impl", "reference", "--grid", "32", "--steps", "5", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = run_bench("--grid", "32", "--steps", "3", "--warmup", "3")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
