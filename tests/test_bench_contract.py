"""The bench contract, checked without a GPU: the committed evidence line (profiles/r2_bench_n1.json, stdout of `python bench.py`
on one B200) carries every key the driver reads, with consistent values; the reference arm runs here on the host cores and
prints the same shape; the B200 arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"]


def load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_committed_bench_line_has_the_contract_keys():
    d = load("r2_bench_n1.json")
    for k in BASE_KEYS:
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].startswith(d["metric"])        # BASELINE's metric names the matcher rows too (`matching*` keys)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = frames of the step / device time of the step
    frames = d["config"]["frames_per_gpu"]
    assert abs(d["value"] - frames / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    # e2e: host buffers, both copies counted, slower than the resident figure and not a repeat of it
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == frames * 640 * 480 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]
    # roofline: achieved / peak = frac, algorithmic bytes of the dominant kernel, traffic from the ncu capture
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] > 0.9 * k["sm_max_mhz"]
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 0 and d["parity_ok"] is True


def test_committed_rows_of_the_other_configs():
    d = load("r2_bench_n1.json")
    assert d["matching"]["parity_ok"] and d["matching"]["kernel"] == "umma"
    assert d["matching_5b"]["parity_ok"] and d["matching_5b"]["roofline"]["bound"] == "tensor"
    for row in ("stereo_euroc", "stereo_kitti"):
        assert d[row]["parity_ok"] and d[row]["pairs_per_s"] > 0 and d[row]["cpu_baseline"]["value"] > 0
    assert d["single_frame_latency"]["median_ms"] < 1.0
    t = load("r2_bench_textured_row.json")["textured_frames"]
    assert t["parity_ok"] and t["level0_fast_candidates_last_frame"] > 2048       # the dense quad-tree paths were exercised


def test_multi_gpu_lines_are_weak_scaling_aggregates():
    one = load("r2_bench_n1.json")["value"]
    for n in (2, 4, 8):
        d = load("r2_bench_n%d.json" % n)
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert 0.9 * n * one < d["value"] < 1.1 * n * one


@pytest.mark.timeout(900)
def test_reference_arm_runs_on_the_host_and_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=850, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_does_not_fall_back_to_the_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
