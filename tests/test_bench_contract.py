"""CPU: the bench lines committed under profiles/ carry every key of the bench.py contract (the driver parses the same
line from a live run), and bench.py itself still spells those keys."""
import json
import os

import pytest

from conftest import ROOT

BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"]


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not recorded yet")
    return json.loads(open(path).read().strip().splitlines()[-1])


def _check_roofline(r):
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
    assert "traffic" in r


@pytest.mark.parametrize("name", ["r1_bench_1gpu.json", "r2_bench_1gpu.json"])
def test_single_gpu_line(name):
    d = _line(name)
    for k in BASE + ["cpu_baseline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("kitti00_shaped_4541kf") and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    _check_roofline(d["roofline"])
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 6e9 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"] and e["results_equal_device_leg"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["sample"] and c["value"] > 0
    assert c.get("loop_ids_and_yaws_equal_gpu", True) is True
    cl = d["clocks"]
    assert cl["sm_mhz"] and cl["sm_max_mhz"] and not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.parametrize("n", [2, 4, 8])
def test_round2_multi_gpu_lines_carry_the_other_configs_and_the_reference_check(n):
    d = _line(f"r2_bench_{n}gpu.json")
    for k in BASE + ["cpu_baseline", "exhaustive_100k", "config3_40k_k50", "config5_40x120_flipped"]:
        assert k in d, k
    assert d["n_gpus"] == n and d["e2e"]["results_equal_device_leg"]
    assert d["cpu_baseline"]["loop_ids_and_yaws_equal_reference"] is True
    assert d["exhaustive_100k"]["winners_equal_oracle_on_3000_entry_prefix"] is True
    assert d["config3_40k_k50"]["first_24_equal_reference"] is True and d["config5_40x120_flipped"]["finds_reversed_revisit"] is True
    _check_roofline(d["roofline"])


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines(n):
    d = _line(f"r1_bench_{n}gpu.json")
    for k in BASE:
        if k == "roofline" and "roofline" not in d:
            pytest.skip("recorded before the multi-GPU line carried a roofline object")
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] in ("weak", "strong") and d["e2e"]["results_equal_device_leg"]
    if "roofline" in d:
        _check_roofline(d["roofline"])


def test_reference_arm_line():
    d = _line("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == _line("r1_bench_1gpu.json")["metric"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]


def test_bench_source_spells_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for k in BASE + ["cpu_baseline", "h2d_bytes_per_step", "d2h_bytes_per_step", "algorithmic_bytes_per_launch", "sm_max_mhz"]:
        assert f'"{k}"' in src, k
