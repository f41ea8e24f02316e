"""bench.py host logic without a GPU: the reference arm prints one JSON line with the contract's keys (on a small
volume), under a fake multi-rank environment only rank 0 prints, and the clock sampler parses nvidia-smi rows."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=e, timeout=300)


import pytest


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_line(kind):
    """kind = reference: the unmodified package staged under baseline/_ref (or /root/reference) runs the steps;
    kind = port: it is hidden, the oracle's PyTorch-eager port stands in."""
    import baseline
    if kind == "reference" and baseline.reference_path() is None:
        pytest.skip("reference package not staged here")
    r = run_bench("--impl", "reference", "--size", "64", "--steps", "2", "--warmup", "1",
                  env={"TAUB_NO_REFERENCE": "1" if kind == "port" else "0"})
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "stencil_sweep_throughput" and d["unit"] == "GLUPS"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    r = run_bench("--impl", "reference", "--gpus", "2", "--size", "64", "--steps", "1", "--warmup", "1", env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = "0"
    r = run_bench("--impl", "reference", "--gpus", "2", "--size", "64", "--steps", "1", "--warmup", "1", env=env)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and "x-slab" in d["config"]["workload"]


def test_clock_sampler_parses_rows():
    import time
    import bench
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    t = time.perf_counter()
    s.rows = [(t, ["0", "1920", "1965", "900.1", "0x4", "Not Active", "Not Active", "Not Active", "Active"]),
              (t + 0.1, ["0", "1935", "1965", "910.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"]),
              (t + 50.0, ["0", "300", "1965", "100.0", "0x0", "Active", "Not Active", "Not Active", "Not Active"])]
    out = s.stop(t - 0.01, t + 0.2)
    assert out["sm_mhz"] == 1927.5 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 2
    assert out["reasons"] == ["sw_power_cap"]          # the idle sample after the timed region is not counted
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]
