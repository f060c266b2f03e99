"""The bench.py JSON contract (driver-facing), checked on the committed bench lines of the last GPU runs
(profiles/r02_bench_n1_final.json, r02_bench_n2_final.json, r02_bench_reference_arm.json): every key the driver and
the judge read is present and self-consistent.  No GPU, no imports from the product."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline")


@pytest.mark.parametrize("name,n", [("r02_bench_n1_final.json", 1), ("r02_bench_n2_final.json", 2)])
def test_own_arm_line(name, n):
    d = _line(name)
    for k in BASE_KEYS[:-1] + ("clocks", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == n and d["metric"] == "pdhg_iterations_per_second" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f32" and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["steps"] > 0
    assert abs(d["value"] - 1.0 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]      # one step = one PDHG iteration
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                  # host copies inside the timed region
    assert d["gpu_launches"] >= d["steps"]               # at least one of our kernels per iteration
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        ref = d["reference_cuda"]
        assert ref["parity_ok"] and max(ref["max_rel_diff"].values()) <= 1e-5


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    for k in BASE_KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "pdhg_iterations_per_second"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_same_iterates_on_one_and_two_gpus():
    a, b = _line("r02_bench_n1_final.json"), _line("r02_bench_n2_final.json")
    assert a["iterate_hash"] == b["iterate_hash"]
