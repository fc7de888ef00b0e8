"""The JSON line bench.py prints is a contract with the driver: the committed lines of the final build
(profiles/bench_r02_n*.json, profiles/bench_r02_reference_arm_n*.json) must carry every key it reads,
with consistent values.  CPU-only: it checks the recorded lines, not a new run."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = [f"bench_r02_n{n}.json" for n in (1, 2, 4, 8)]


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


@pytest.mark.parametrize("name", LINES)
def test_own_arm_line_has_the_contract_keys(name):
    d = _load(name)
    n = int(name.split("_n")[1].split(".")[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "chunks/sec (gen+light+mesh)" and d["unit"] == "chunks/s"
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["vs_baseline"] is None  # BASELINE.md holds no published number for this metric
    assert d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "c3_128x128_gen_light_mesh" and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    # value = units of all ranks / device time of the K steps
    assert d["value"] == pytest.approx(16384 * 1e3 / d["ms_per_step"], rel=1e-6)
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert e["unit"] == "chunks/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]  # host buffers and copies inside the timed region: never the device number
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor", "fp64") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0 < r["frac"] < 1
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        assert r["traffic"] and r["traffic"] > 0
        # the kernels' own times add up to the step: nothing else hides in the timed region
        assert d["kernel_ms_sum"] <= d["ms_per_step"] < 1.05 * d["kernel_ms_sum"]


@pytest.mark.parametrize("name", ["bench_r02_reference_arm_n1.json", "bench_r02_reference_arm_n8.json"])
def test_reference_arm_line(name):
    d, own = _load(name), _load("bench_r02_n1.json")
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better", "dtype"):
        assert d[k] == own[k]
    assert d["config"]["workload"] == own["config"]["workload"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    b = d["cpu_baseline"]
    assert b["value"] == d["value"] and b["cores"] >= 1 and b["kind"] in ("port", "reference")
    assert own["e2e"]["value"] > 100 * d["value"]  # the headline ratio, orders of magnitude
