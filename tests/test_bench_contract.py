"""The bench line contract (no GPU): the committed lines of the last GPU visit carry every key the driver reads, with
consistent values.  The lines were printed by `python bench.py` / `python bench.py --impl reference` on a B200
(profiles/r01_bench_C2_s8c.json, r01_bench_ref_C2_s8a.json)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_own_arm_line():
    d = _line("r01_bench_C2_s8c.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "apg_iterations_per_sec" and d["unit"] == "iter/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert "larger than L2" in d["config"]["cold_cache"]
    iters = d["config"]["iterations_per_solve"]
    assert abs(d["value"] - iters / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] <= d["value"] * 1.001          # same metric through the host-buffer call, not a copy
    assert e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == d["unit"] and b["sample"]
    # the reformulations ride along, never as the headline
    for k in ("alt_formulation", "alt_formulation_shared"):
        assert d[k]["unit"] == d["unit"] and d[k]["roofline"]["algorithmic_bytes_per_launch"] < r["algorithmic_bytes_per_launch"]
    assert d["config"]["factors"] == "full"


def test_reference_arm_line():
    d = _line("r01_bench_ref_C2_s8a.json")
    assert d["impl"] == "reference" and d["metric"] == "apg_iterations_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    b = d["cpu_baseline"]
    assert b["value"] == d["value"] and b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["sample"]


def test_same_workload_in_both_arms():
    a, b = _line("r01_bench_C2_s8c.json"), _line("r01_bench_ref_C2_s8a.json")
    assert a["config"]["workload"] == b["config"]["workload"]
    assert a["config"]["iterations_per_solve"] == b["config"]["iterations_per_solve"]
