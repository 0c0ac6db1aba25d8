"""The bench.py JSON contract, checked on the lines recorded on the B200 boxes (profiles/r01_final_*):
every key the driver reads is present with the right type and the numbers are internally consistent.
(The lines themselves were produced by `bench.py` on GPU; this guards the schema on CPU.)"""
import json
import os

import pytest

from conftest import ROOT

PROFILES = os.path.join(ROOT, "profiles")
REQUIRED = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int,
            "ms_per_step": float, "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict,
            "e2e": dict, "gpu_launches": int, "roofline": dict, "clocks": dict}


def load(name):
    p = os.path.join(PROFILES, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not recorded")
    return json.loads([l for l in open(p) if l.startswith("{")][-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_our_arm_line(n):
    j = load(f"r01_final_bench_n{n}.json")
    for k, t in REQUIRED.items():
        assert k in j, k
        assert isinstance(j[k], t) or (t is float and isinstance(j[k], int)), (k, type(j[k]))
    assert j["metric"] == "nvlink_probe_ms" and j["unit"] == "ms" and j["higher_is_better"] is False
    assert j["n_gpus"] == n and j["vs_baseline"] is None and j["data"] == "synthetic" and j["scaling"] == "weak"
    assert j["warmup"] >= 3 and j["gpu_launches"] == j["steps"] * n
    assert "workload" in j["config"] and "model" not in j["config"]
    e = j["e2e"]
    assert e["unit"] == "ms" and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = j["roofline"]
    assert set(["bound", "achieved", "peak", "unit", "frac", "traffic"]) <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    c = j["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and isinstance(c["reasons"], list)
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert j["reachability_all_ones"] is True and j["verdict"] is True
    # the wall clock of the timed loop agrees with the per-call figure (no work hidden outside the loop)
    assert abs(j["ms_per_step"] - e["value"]) / e["value"] < 0.05
    if n == 1:
        assert r["bound"] == "hbm" and 0.5 < r["frac"] < 1.1 and r["traffic"] >= r["algorithmic_bytes_per_launch"]
        cb = j["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    else:
        assert r["bound"] == "nvlink" and 0.5 < r["frac"] < 1.0
        nv = j["nvlink_counters"]
        assert abs(nv["tx_kib_delta"] / nv["algorithmic_kib_per_direction"] - 1) < 1e-3  # counters == algorithmic bytes
        assert j["per_link_gbps"]["run_to_run_spread_read"] < 0.02 and j["per_link_gbps"]["run_to_run_spread_write"] < 0.02
        assert j["value"] < 5.0  # north_star: < 5 ms


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_reference_arm_line(n):
    j = load(f"r01_final_ref_n{n}.json")
    assert j["impl"] == "reference" and j["metric"] == "nvlink_probe_ms" and j["unit"] == "ms"
    assert j["higher_is_better"] is False and j["n_gpus"] == n and j["gpu_launches"] == 0 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ours = load(f"r01_final_bench_n{n}.json")
    if "workload" in j["config"] and "reference_path" in j["config"]:
        assert j["config"]["workload"] == ours["config"]["workload"]  # both arms answer the same question
