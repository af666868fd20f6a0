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
    # the CPU arm plans with the Python A* and runs the oracle: the product's CUDA library is never mapped
    assert d["native_libraries_loaded"] == ["libobca_oracle.so"]
    assert sum(d["status_histogram"].values()) == 64 and sum(d["iters_histogram"].values()) == 64
    assert d["config"]["cfg"] == 3 and d["config"]["batch_per_gpu"] == 8192


def test_reference_arm_other_configurations():
    """--cfg selects the other single-launch BASELINE configurations; --start reference the reference's own start"""
    d = _run(["--impl", "reference", "--cfg", "5", "--batch", "48", "--steps", "1", "--warmup", "0"])
    assert d["config"]["cfg"] == 5 and d["config"]["mode"].startswith("FIXED_SET") and d["config"]["rows"] == 24
    assert "cfg 5" in d["metric"] and sum(d["status_histogram"].values()) == 48
    d = _run(["--impl", "reference", "--cfg", "2", "--batch", "48", "--steps", "1", "--warmup", "0", "--start", "reference"])
    assert d["config"]["cfg"] == 2 and "zeros" in d["config"]["init"] and d["success_rate"] > 0.9


def test_reference_arm_other_ranks_are_silent():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line_small_batch():
    d = _run(["--batch", "512", "--steps", "3", "--warmup", "3", "--cpu-sample", "128"])
    assert BASE_KEYS | {"roofline", "roofline_fp64", "clocks", "gpu_launches", "status_histogram", "iters_histogram"} <= set(d)
    # three launches per step: the first-pass kernel, the recovery block beside it, the recovery over what is left
    assert d["gpu_launches"] == 9 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["roofline_fp64"]["peak"] > 10 and d["roofline_fp64"]["unit"] == "TFLOP/s"
    assert sum(d["status_histogram"].values()) == 512
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
