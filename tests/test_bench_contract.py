"""The committed bench lines under profiles/ carry every key the bench contract names (CPU check of the recorded
evidence; bench.py itself needs a GPU), and the reference arm describes the same workload as the GPU arm."""
import glob
import json
import os

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config")


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "profiles", "r2_bench_n*.json"))))
def test_gpu_arm_lines(name):
    d = _load(name)
    for k in BASE + ("roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "share_msm_g1_throughput" and d["unit"] == "Mpts/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    n = d["n_gpus"]
    assert abs(d["value"] - d["config"]["total_points"] / d["ms_per_step"] / 1e3) / d["value"] < 1e-6
    assert d["config"]["total_points"] == n * d["config"]["points_per_gpu"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    if n == 1:
        cb = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in cb, k
        assert cb["kind"] in ("port", "reference") and cb["unit"] == d["unit"]
        for shape, v in d["extra"]["prove"].items():
            if "matches_cpu_proof" in v:
                assert v["matches_cpu_proof"] is True, shape
    else:
        assert "strong" in d["extra"] and d["extra"]["strong"].get("msm_2p24", {}).get("same_result_as_1gpu") is True


def test_reference_arm_line():
    d, g = _load("r2_bench_reference_arm.json"), _load("r2_bench_n1.json")
    for k in BASE + ("impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better", "config"):
        assert d[k] == g[k], k                                  # same workload description as the GPU arm
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
