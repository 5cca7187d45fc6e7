"""bench.py contract: one JSON line with the keys the driver reads (metric/value/unit/n_gpus/steps/warmup/ms_per_step/..., `roofline`,
`cpu_baseline`, `e2e`, `clocks`, `gpu_launches`), for the engine arm (GPU) and the reference arm (CPU oracle, no GPU needed)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"}


def run_bench(args, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "candidates scored/sec" and d["unit"] == "candidates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_engine_arm_line(capi):
    d = run_bench(["--steps", "2", "--warmup", "3", "--batch-log2", "26", "--cpu-seconds", "1"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 1e9 and d["scaling"] == "weak"
    assert d["gpu_launches"] == 2 * 2  # sweep + final kernel per step
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 2 and r["peak"] > 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 336 and e["d2h_bytes_per_step"] == 24
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    assert d["best"]["score"] < 17.86  # the sweep improves on Winograd's growth factor (17.853)
