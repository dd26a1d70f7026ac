"""bench.py's CPU side of the contract (no GPU needed): the reference arm prints one JSON line with the agreed keys for
the default workload and for the extra full-size workloads, ranks other than 0 stay silent, and the product arm refuses
to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.parametrize("workload,frames", [("kitti", "4"), ("tum", "4")])
def test_reference_arm_line(workload, frames):
    pytest.importorskip("cv2")
    r = run(["--impl", "reference", "--workload", workload, "--frames", frames, "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["unit"] == "frames/s" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u8"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    if workload == "kitti":
        assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"].split(" (")[0] or "1241x376" in d["metric"]


def test_reference_arm_other_ranks_are_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--frames", "4", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run(["--frames", "4", "--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_extra_workloads_are_orb_only():
    r = run(["--workload", "4k", "--mode", "reference"])
    assert r.returncode != 0 and "ORB-mode" in r.stderr
