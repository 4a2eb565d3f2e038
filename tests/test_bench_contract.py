"""The bench.py JSON contract, checked on the committed B200 lines under profiles/ (no GPU needed): every key the
driver and the judge read is present and self-consistent."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


@pytest.mark.parametrize("name,n", [("r1_bench_1gpu.json", 1), ("r1_bench_2gpu.json", 2), ("r1_bench_4gpu.json", 4),
                                    ("r1_bench_8gpu.json", 8)])
def test_b200_arm_line(name, n):
    d = _line(name)
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["n_gpus"] == n and d["steps"] >= 1 and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None          # BASELINE.md holds no published number for this metric
    assert "frames" in d["unit"] and d["metric"].startswith("forecast_frames")
    assert "forecast frames/sec" in base["metric"]
    assert "workload" in d["config"] and "model" not in d["config"]
    # value is the whole-job aggregate: 12 forecast frames x 32 sequences x N GPUs per step
    frames = 12 * d["config"]["batch_per_gpu"] * n
    assert d["value"] == pytest.approx(frames / (d["ms_per_step"] * 1e-3), rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["value"] > 0 and e["value"] != d["value"]
    assert e["h2d_bytes_per_step"] == d["config"]["batch_per_gpu"] * 384 * 384 * 25 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 1000
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0.5 < r["frac"] <= 1.0
    assert "traffic" in r
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == d["unit"] and cb["sample"]
        assert r["traffic"] is None or r["traffic"] > 0


def test_reference_arm_line():
    d, mine = _line("r1_bench_reference_arm.json"), _line("r1_bench_1gpu.json")
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["metric"] == mine["metric"] and d["unit"] == mine["unit"] and d["higher_is_better"] is True
    assert d["config"]["workload"] == mine["config"]["workload"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert mine["e2e"]["value"] / d["value"] > 100      # the headline ratio the driver computes


def test_round2_line_has_honest_roofline_and_secondary_passes():
    d = _line("r2_bench_1gpu.json")
    r = d["roofline"]
    assert 0.5 < r["frac_executed"] < r["frac"] <= 1.0          # executed MMA work < nominal (sub-pixel upsample)
    assert r["executed_tflops"] == pytest.approx(r["frac_executed"] * r["peak"], rel=1e-9)
    assert r["step_frac_of_peak_executed"] < r["step_frac_of_peak"]
    for key in ("n256", "n128"):
        tr = r["traffic_ncu"][key]
        assert os.path.exists(os.path.join(ROOT, tr["source"])) and tr["dram_bytes"] > 1e8
        assert 40 < tr["tensor_pipe_pct_elapsed"] < 100
    assert r["traffic"] == r["traffic_ncu"]["n256"]["dram_bytes"]
    sec = r["secondary"]
    assert set(sec) == {"stage_vil", "predict_linear", "metrics"}
    for name, s in sec.items():
        assert s["bound"] == "hbm" and 0 < s["frac"] < 1.05 and s["gbs"] == pytest.approx(s["frac"] * r["hbm_peak_gbs"], rel=1e-9)
    assert sec["stage_vil"]["algorithmic_bytes_per_launch"] == 32 * 25 * 384 * 384 * 5
    assert sec["predict_linear"]["algorithmic_bytes_per_launch"] == 32 * 921600
    assert sec["metrics"]["algorithmic_bytes_per_launch"] == 32 * 12 * 1179648
    assert d["parity_check"] is None                                   # single rank
    ref = _line("r2_bench_reference_arm.json")
    assert ref["config"] == d["config"] and ref["impl"] == "reference"


def test_round2_two_gpu_line_checks_the_reduction_on_device():
    d = _line("r2_bench_2gpu.json")
    assert d["n_gpus"] == 2 and d["parity_check"] == "ok"
    assert d["value"] == pytest.approx(2 * 12 * 32 / (d["ms_per_step"] * 1e-3), rel=1e-6)


@pytest.mark.parametrize("cfg", ["posaware", "vit", "disc"])
def test_round2_aux_config_lines(cfg):
    d = _line(f"r2_bench_config_{cfg}.json")
    assert d["unit"] == "frames/s" and d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    assert "BASELINE configs[" in d["config"]["workload"] and "model" not in d["config"]
    assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["step_frac_of_peak"] < 1
