"""The bench lines committed under profiles/ carry every key of the bench.py contract (checked on CPU; the lines themselves were
produced on B200 by the commands named in profiles/README.md)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "profiles", "bench_r02_*gpu_*.json")))
                         + ["bench_r02_c1.json", "bench_r02_c2.json", "bench_r02_c3.json", "bench_r02_c4.json"])
def test_own_arm_lines(name):
    d = _load(name)
    for k in BASE_KEYS:
        assert k in d, (name, k)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].startswith(d["metric"]) and d["unit"] == "voxel-iterations/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["vs_baseline"] is None
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["roofline"] is not None:
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert 0 < r["frac"] < 1.05
    if d["e2e"] is not None:
        e = d["e2e"]
        # (c1, 64^3, is launch-bound: its kernel-only figure carries the per-kernel event pairs of the timed region and ends up
        # below the end-to-end figure, which runs without them)
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < (1.5 if "c1" in name else 1.0) * d["value"]
    if d["n_gpus"] > 1:
        assert d["scaling"] in ("weak", "strong") and d["nvlink"]["achieved_gbs_per_direction"] < 900


def test_headline_line():
    d = _load("bench_r02_c2.json")
    assert d["n_gpus"] == 1 and d["config"]["grid"] == [256, 256, 256]
    assert d["parity_256"]["ok"] and d["parity_256"]["max_rel"] <= 1e-10
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and "256x256x256" in c["sample"]
    # throughput, time per step and the roofline fraction of the whole iteration agree with each other
    assert abs(d["value"] - 256 ** 3 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
    h = d["iteration_hbm"]
    assert abs(h["frac_of_peak"] - 728.0 * d["value"] / 1e9 / h["peak_gbs"]) < 1e-9


def test_reference_arm_line():
    d = _load("bench_r02_reference_arm.json")
    assert d["impl"] == "reference" and d["config"]["same_config"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"]


def test_strong_scaling_runs_are_the_same_solve():
    runs = [_load("bench_r02_%dgpu_strong1024.json" % n) for n in (2, 4, 8)]
    one = _load("bench_r02_c2.json")["e2e"]
    for d in runs:
        s = d["strong_scaling"]
        assert s["iterations"] == one["iterations"] == 52
        assert abs(s["final_residual"] - one["final_residual"]) <= 1e-10 * one["final_residual"]
        assert abs(s["mean_stress"][0] - one["mean_stress_11"]) <= 1e-9 * abs(one["mean_stress_11"])
