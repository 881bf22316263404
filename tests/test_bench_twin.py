"""tools/perceive_bench.cpp — the compiled twin of the `perceive bench` subcommand
(rust/perceive-cli/cmd/bench.rs, unbuilt: no Rust toolchain) over the same C ABI.  Built with g++, run on
BASELINE config 1 and held to the Python binding: same hits for the same step, same workload description
as bench.py, a sane throughput figure."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def bench_exe(pcv_lib, tmp_path_factory):
    from perceive_b200 import _ffi
    exe = tmp_path_factory.mktemp("bench") / "perceive_bench"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "tools" / "perceive_bench.cpp"), "-o", str(exe),
           "-L", str(_ffi.LIB_PATH.parent), "-lperceive_cuda", f"-Wl,-rpath,{_ffi.LIB_PATH.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_bench_twin_builds_and_rejects_bad_arguments(bench_exe):
    r = subprocess.run([str(bench_exe), "--config", "c9"], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown workload" in r.stderr
    import torch
    if not torch.cuda.is_available():  # no device: a loud failure, never a CPU path
        r = subprocess.run([str(bench_exe), "--config", "c1"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_bench_twin_on_config1_matches_the_python_binding(bench_exe, orc):
    import perceive_b200 as pb
    sys.path.insert(0, str(ROOT))
    import bench as pybench
    steps, warmup = 30, 5
    r = subprocess.run([str(bench_exe), "--config", "c1", "--steps", str(steps), "--warmup", str(warmup), "--dump-ids"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    w = pybench.WORKLOADS["c1"]
    assert out["config"]["workload"] == w["text"] and out["config"]["rows"] == w["rows"] and out["config"]["k"] == w["k"]
    assert out["parity"]["planted_top1"] is True and out["steps"] == steps and out["gpu_launches"] == steps
    # the last timed step used query (warmup + steps - 1) % pool of the query stream (seed 2)
    pool = steps + warmup  # 64 MB / 1.5 KB is far more than 35 batches
    q = orc.synth_rows(pybench.QUERY_SEED, 0, (warmup + steps - 1) % pool, 1, w["dim"])
    with pb.Index(w["dim"]) as ix:
        ix.generate_synthetic(w["rows"], pybench.CORPUS_SEED)
        ids, scores, _, _ = ix.search(q, w["k"])
    assert out["last_step_ids"] == ids[0].tolist()
    assert np.array_equal(np.array(out["last_step_score_bits"], dtype=np.uint32).view(np.float32), scores[0])
    # a 15 MB corpus is L2-resident: anywhere from ~10 us (device) to a few hundred us (host round trip) per query
    assert 2e3 < out["e2e"]["value"] < 2e5 and out["value"] >= out["e2e"]["value"]
