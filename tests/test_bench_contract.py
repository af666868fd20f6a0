"""bench.py contract: one JSON line with the keys the driver reads.  CPU: the reference arm (oracle on host cores);
GPU (-m gpu): the B200 arm on a small batch, plus __graft_entry__.smoke()."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "64"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_are_silent():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line_small_batch():
    d = _run(["--batch", "512", "--steps", "3", "--warmup", "3", "--cpu-sample", "128"])
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(d)
    assert d["gpu_launches"] == 3 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["bytes_per_solve"] == 7160
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["success_rate"] > 0.9


@pytest.mark.gpu
def test_graft_entry_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.smoke()
