"""The bench.py JSON line (driver contract) — checked on the committed round evidence, which are
verbatim lines printed by `python bench.py` on a B200, and on bench.py's own workload table."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")

OURS = ["r01_bench_c2.json", "r01_bench_c3.json", "r01_bench_c1.json", "r01_bench_c2_8gpu.json", "r01_bench_c4_8gpu.json",
        "r02_bench_default.json", "r02_bench_2gpu.json", "r02_bench_4gpu.json", "r02_bench_8gpu.json"]
REFERENCE = ["r01_bench_c2_reference.json", "r01_bench_c3_reference.json", "r02_bench_reference.json"]


def _line(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", OURS)
def test_our_arm_line_carries_every_contract_key(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "queries/sec" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert abs(e["value"] - d["value"]) > 1e-6           # measured separately, not a copy of `value`
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


@pytest.mark.parametrize("name", REFERENCE)
def test_reference_arm_line(name):
    d = _line(name)
    assert d["impl"] == "reference" and d["metric"] == "queries/sec" and d["unit"] == "queries/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")


def test_workload_table_names_the_baseline_configs():
    import sys
    sys.path.insert(0, ROOT)
    import bench

    w = bench.WORKLOADS
    assert w["c2"][:5] == (1_000_000, 768, 2, 10_000, 10) and w["c2"][5] == "flat"     # configs[1], the default
    assert w["c3"][:5] == (1_000_000, 128, 1, 10_000, 10) and w["c3"][5] == "hnsw"     # configs[2]
    assert w["c1"][:5] == (100_000, 128, 1, 1_000, 10) and w["c1"][5] == "hnsw"        # configs[0]
    assert w["c4"][:5] == (10_000_000, 768, 3, 10_000, 10) and w["c4"][5] == "flat"    # configs[3]


@pytest.mark.parametrize("name", ["r02_bench_default.json", "r02_bench_2gpu.json", "r02_bench_4gpu.json", "r02_bench_8gpu.json"])
def test_round2_lines_are_verified_against_the_oracle_and_carry_their_secondary_blocks(name):
    d = _line(name)
    v = d["verified"]
    assert v["identical"] is True and v["identical_pageable_call"] is True and v["queries"] >= 8
    assert d["e2e"]["pageable"]["value"] > 0
    r = d["roofline"]
    assert abs(r["frac_sustained"] - r["achieved"] / 1386.4) < 1e-3 and r["frac_burst"] < r["frac_sustained"]
    names = [b["workload"].split(":")[0] for b in d["secondary"]]
    assert names == (["c3", "build"] if d["n_gpus"] == 1 else ["c4"])
    for b in d["secondary"]:
        assert b["verified"]["identical"] is True
        if b["workload"].startswith("c3"):
            assert b["verified"]["recall_at_k_gpu"] == b["verified"]["recall_at_k_oracle"]
            assert all(p["identical_to_oracle_sample"] for p in b["ef_sweep"]) and len(b["ef_sweep"]) >= 4
            assert 0 < b["roofline"]["frac_dram"] < b["roofline"]["frac"]


def test_traffic_table_points_at_committed_captures():
    # roofline.traffic is read from profiles/traffic.json: every entry names the ncu summary it was taken from
    with open(os.path.join(PROFILES, "traffic.json")) as f:
        table = json.load(f)
    assert table
    for key, entry in table.items():
        assert entry["bytes"] > 0
        src = entry["source"].split(" ")[0]
        assert src.startswith("profiles/") and os.path.exists(os.path.join(ROOT, src)), (key, src)
