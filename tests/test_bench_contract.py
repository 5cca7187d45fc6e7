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
    assert "candidates_per_gpu_per_step" in d["config"]  # same key names as the engine arm's config
    cfg = d["configs"]  # the other BASELINE configs, CPU oracle on the same inputs
    assert cfg["C4_orbit_3x4x7_nnz"]["value"] > 0 and cfg["C4_orbit_3x4x7_G2"]["value"] > 0
    assert cfg["C3_sparsifier_4x4x4"]["search_1core"]["cores"] == 1 and cfg["C3_sparsifier_4x4x4"]["pipeline_c11"]["value"] > 0
    assert cfg["C5_mmcheck_32x32x32"]["all_correct"] and cfg["C5_mmcheck_32x32x32"]["unit"] == "samples/s"


@pytest.mark.gpu
def test_engine_arm_line(capi):
    d = run_bench(["--steps", "2", "--warmup", "3", "--batch-log2", "26", "--cpu-seconds", "1"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 1e9 and d["scaling"] == "weak"
    assert d["gpu_launches"] == 2 * 3  # sweep + final + slot-pack kernel per step
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1 and r["peak"] > 1
    assert r["kernel"] == "orbit_sweep8x_kernel" and r["lanes_per_imad"] == 4  # frac is stated against lanes x the measured IMAD peak
    cfg = d["configs"]
    for key in ("C4_orbit_3x4x7_nnz", "C4_orbit_3x4x7_G2"):
        e = cfg[key]
        assert e["value"] > 1e8 and e["roofline"]["lanes_per_imad"] == 4 and 0 < e["roofline"]["frac"] < 1 and e["e2e"]["value"] > 0 and e["cpu_baseline"]["value"] > 0
    c3 = cfg["C3_sparsifier_4x4x4"]
    assert c3["search_c128"]["value"] > 1e10 and c3["search_c128"]["e2e"]["value"] > 1e10 and c3["pipeline_c11"]["consistent"]
    assert c3["pipeline_c11"]["value"] > 10 * c3["pipeline_c11"]["round1_value"] / 2
    c5 = cfg["C5_mmcheck_32x32x32"]
    assert c5["batch_4096"]["all_samples_agree"] and c5["batch_32"]["all_samples_agree"] and c5["batch_4096"]["value"] > 1e5
    assert cfg["C2_strong"]["scaling"] == "strong" and cfg["C2_strong"]["value"] > 1e9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 336 and e["d2h_bytes_per_step"] == 24
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    assert d["best"]["score"] < 17.86  # the sweep improves on Winograd's growth factor (17.853)
