"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the
same metric / unit / config as our arm, a cpu_baseline describing the run and a zero-copy e2e block; the cpu_baseline
legs of our arm (a process of their own) cover every row with both code paths; our arm refuses to run without a device."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout)


def test_reference_arm_line():
    sys.path.insert(0, str(ROOT))
    import bench
    r = _run("--impl", "reference", "--bases", "3000000", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["config"] == bench.workload_config(3000000)          # the same config object as our arm's
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "whole workload" in cb["sample"]
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] * 1e9 - 2 * 3000000) < 1e-3 * 2 * 3000000   # value = 2 n / step time


def test_reference_arm_other_ranks_exit_silently():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--bases", "1000000", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_suite_covers_every_row():
    r = _run("--cpu-suite", "--bases", "8000000")
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for name in ("codec_full", "codec", "kmers", "hdist", "base_counts", "encode_batch", "cfg0"):
        assert d[name] and all(row["value"] > 0 and row["best"] >= row["value"] >= row["worst"] and row["reps"] >= 2 for row in d[name]), name
    isas = {row["isa"] for row in d["codec"]}
    assert "scalar" in isas                                        # the nosimd analogue is always timed
    assert {row["cores"] for row in d["codec"]} >= {1}
    assert d["host"]["threads"] >= 1 and "march=native" in d["host"]["compiler"]


def test_our_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1", "--bases", "1000000", timeout=300)
    assert r.returncode != 0 and "CUDA device" in (r.stdout + r.stderr)
