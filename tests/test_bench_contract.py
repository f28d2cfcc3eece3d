"""The JSON lines bench.py prints: the reference arm is run here (it needs no GPU), the GPU arm is checked on the committed
records of the final builds (profiles/r1_bench_default.json, profiles/r2_bench_default.json, profiles/r2_bench_dense.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"]


def check_common(d):
    for k in BASE_KEYS:
        assert k in d, k
    assert d["unit"] == "scans/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    assert d["value"] > 0 and d["ms_per_step"] > 0
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-frames", "8"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    check_common(d)
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"]


def test_committed_gpu_record():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_default.json")))
    check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cl = d["clocks"]
    assert cl["sm_mhz"] > 0 and cl["sm_max_mhz"] >= cl["sm_mhz"] and not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert set(d["stage_ms"]) == {"extract", "scan_ds", "assoc_solve", "map_update"}


def test_committed_round2_records():
    """Round 2 added: steady-state value after a preroll + the young-map value, the latency block, the copy-only ceilings, the
    issue roofline; and the configs[2] workload as a second line."""
    for name, min_map in (("r2_bench_default.json", 5e4), ("r2_bench_dense.json", 1e6)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        check_common(d)
        assert d["config"]["preroll_frames"] >= 300 and d["config"]["mean_map_points"] >= min_map
        assert d["young_map"]["value"] > 0 and d["latency"]["gpu_ms_per_frame"] > 0 and d["latency"]["cpu_one_core_ms_per_frame"] > d["latency"]["gpu_ms_per_frame"]
        e = d["e2e"]
        assert e["h2d_copy_only_per_scan_gbs"] > 0 and e["h2d_copy_only_per_group_gbs"] > 0 and 0 < e["e2e_efficiency"] <= 1.05
        r = d["roofline"]
        assert r["traffic"] is not None and "NOT measured in this run" in r["traffic_source"]
        assert 0 < r["issue"]["frac"] < 1 and r["issue"]["warp_instructions_per_launch"] > 0
        q = d["large_map"]["knn_query"]
        assert q["queries_per_s"] > 9e8 and 0 < q["issue_roofline"]["frac"] < 1
        assert d["gpu_launches"] > 0 and d["cpu_baseline"]["kind"] == "port"
