"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm runs end to end on the
CPU and prints the agreed keys; both arms build their `config` from one function; `roofline.traffic` comes from the
committed ncu summaries; the clock sampler's windowing."""
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_reference_arm_prints_the_agreed_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "extrapolated" not in cb  # config 1 is timed in full
    assert d["config"] == bench.config_of(dict(bench.WORKLOADS["c1"]), 1)


def test_batched_reference_arm_reports_what_was_timed():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c4", "--rows", "40000000",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    cb = d["cpu_baseline"]
    assert cb["sample_rows"] < 40_000_000 and "SAMPLE" in cb["sample"]
    ex = cb["extrapolated"]
    assert abs(ex["value"] - d["value"] * cb["sample_rows"] / 40_000_000) < 1e-6 * d["value"]
    assert d["config"]["rows"] == 40_000_000  # the workload, not the sample


def test_both_arms_share_one_config_function():
    src = (ROOT / "bench.py").read_text()
    assert src.count('"config": config_of(') + src.count("cfg = config_of(") == 2  # the reference line and the GPU record
    for name, w in bench.WORKLOADS.items():
        for world in (1, 8):
            c = bench.config_of(dict(w), world)
            assert c["workload"] == w["text"] and c["rows"] == w["rows"] and c["n_gpus"] == world
            assert ("row" in c["sharding"]) == (world > 1)


def test_traffic_comes_from_the_committed_ncu_summaries():
    t = bench.ncu_traffic("c2", 1_000_000)
    assert t and abs(t["bytes"] / 1.536e9 - 1.0) < 0.01 and "profiles/" in t["source"]
    t = bench.ncu_traffic("c4", 12_500_000)  # what one rank of the 8-GPU run launches
    assert t and 7.9e9 < t["bytes"] < 8.1e9  # 81 273 tiles x 128 rows x 768 B of hi plane + the candidate writes
    assert bench.ncu_traffic("c4", 123) is None and bench.ncu_traffic("c9", 1) is None
    for _, _, files, _, _ in bench.NCU_SUMMARIES:
        assert any((ROOT / f).exists() for f in files), files


def test_bf16_rounding_helper_matches_the_oracle(orc):
    x = np.random.default_rng(0).standard_normal(5000).astype(np.float32)
    assert np.array_equal(bench.bf16_round(x), orc.round_bf16(x))


def test_clock_sampler_window_picks_the_samples_of_a_region():
    s = bench.ClockSampler(0)
    s.proc = object()  # pretend nvidia-smi is running; samples are injected
    now = time.monotonic()
    s.samples = [(now - 1.0, 1965.0, 1965.0, []), (now - 0.5, 1200.0, 1965.0, ["sw_power_cap"]), (now - 0.4, 1300.0, 1965.0, ["sw_power_cap"]),
                 (now + 0.0, 1965.0, 1965.0, [])]
    w = s.window(now - 0.55, now - 0.35)
    assert w["samples"] == 2 and w["sm_mhz"] == 1250.0 and w["reasons"] == ["sw_power_cap"] and w["sm_max_mhz"] == 1965.0
    w = s.window(now - 0.80, now - 0.795)  # a region shorter than the sampling period: the nearest samples
    assert w["samples"] == 2 and w["sm_mhz"] is not None
