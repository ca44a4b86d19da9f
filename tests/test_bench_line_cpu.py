"""CPU: the committed bench line of the final build (profiles/bench_r2_end.json) carries every key of the measurement contract
and its derived numbers are consistent (value = samples / time, roofline.frac = achieved / peak, kernel shares)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_final_line_contract():
    d = _line("bench_r2_end.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    B = d["config"]["batch_per_gpu"]
    assert abs(d["value"] - B / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]          # whole-job samples / max-over-ranks time
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9
    assert abs(r["achieved"] - r["algorithmic_flops"] / (r["us_per_launch"] * 1e-6) / 1e12) <= 1e-6 * r["achieved"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert max(d["joint_err_mm"]) <= 0.05                                                 # north_star's joint bar, on the timed configuration


def test_multi_gpu_lines_are_weak_scaling_of_the_same_workload():
    one = _line("bench_r2_end.json")
    for n in (2, 4):
        d = _line(f"bench_r2_end_{n}gpu.json")
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["config"]["batch_per_gpu"] == one["config"]["batch_per_gpu"]
        assert abs(d["value"] - n * d["config"]["batch_per_gpu"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
        assert len(d["rank_ms_per_step"]) == n and abs(max(d["rank_ms_per_step"]) - d["ms_per_step"]) < 1e-3
        assert d["value"] / (n * one["value"]) > 0.9
