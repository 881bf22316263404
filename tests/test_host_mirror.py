"""include/perceive_search.hpp — the C++ host-side mirror of perceive_core::search (the reference's
host language, Rust, has no toolchain here).  A small C++ program uses it the way a caller of the
reference would; this file builds it, runs it and checks what it printed against the oracle."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from test_searcher_sqlite import DIM, _file_db

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def mirror_exe(pcv_lib, tmp_path_factory):
    from perceive_b200 import _ffi
    exe = tmp_path_factory.mktemp("mirror") / "host_mirror_check"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"),
           str(Path(__file__).parent / "host_mirror_check.cpp"), "-o", str(exe),
           "-L", str(_ffi.LIB_PATH.parent), "-lperceive_cuda", f"-Wl,-rpath,{_ffi.LIB_PATH.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run(exe, *args, env=None):
    import os
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=300,
                       env=None if env is None else dict(os.environ, **env))
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    return {d["step"]: d for d in map(json.loads, r.stdout.splitlines())}


def test_mirror_compiles_and_host_side_behaviour(mirror_exe, orc, tmp_path):
    """Codec, error codes, the empty Searcher — and, on a box without a GPU, that build() on a real
    database fails loudly instead of falling back to a CPU path."""
    import torch
    path, _, _ = _file_db(orc, tmp_path, n=60)
    out = _run(mirror_exe, "cpu", path)
    assert "cpu_ok" in out
    if not torch.cuda.is_available():
        assert "no_device" in out


def _items(step):
    return step["ids"], np.array(step["score_bits"], dtype=np.uint32).view(np.float32)


@pytest.mark.gpu
def test_mirror_searcher_matches_oracle(mirror_exe, orc, tmp_path):
    path, conn, live = _file_db(orc, tmp_path)
    ids = np.array(sorted(live), dtype=np.int64)
    rows = np.stack([live[int(i)][1] for i in ids])
    srcs = np.array([live[int(i)][0] for i in ids], dtype=np.int64)
    q = orc.synth_rows(12, 0, 0, 1, DIM)[0]
    qfile = tmp_path / "query.f32"
    q.astype("<f4").tofile(qfile)

    def want(k, flt=(1, 2, 3), keep=None):
        m = np.ones(len(ids), bool) if keep is None else keep
        w_ids, w_scores, _ = orc.search(rows[m], ids[m], q, k, source_ids=srcs[m], sources=list(flt), mode=orc.MODE_F32_V1)
        return w_ids.tolist(), w_scores

    import sqlite3
    top_ids, _ = want(5)
    victim = top_ids[1]  # hidden in the database further down, before the second run
    out = _run(mirror_exe, "gpu", path, 7, qfile)
    assert out["built"]["dim"] == DIM and out["built"]["n_sources"] == 3
    for step, k, flt in (("all", 10, (1, 2, 3)), ("src2", 10, (2,)), ("batch0", 10, (1, 2, 3)), ("batch1", 10, (1, 2, 3)),
                         ("hidden_ignored", 5, (1, 2, 3)), ("moved", 3, (1, 2, 3))):
        g_ids, g_scores = _items(out[step])
        w_ids, w_scores = want(k, flt)
        assert g_ids == w_ids, step
        assert np.array_equal(g_scores, w_scores), step  # bit-identical reference distances
    assert out["none"]["ids"] == [] and out["unknown"]["ids"] == []
    hidden = [top_ids[0], top_ids[2]]
    g_ids, g_scores = _items(out["hidden_filtered"])
    w_ids, w_scores = want(5, keep=~np.isin(ids, hidden))
    assert g_ids == w_ids and np.array_equal(g_scores, w_scores)
    assert out["like"]["ids"][0] == top_ids[0] and out["like_missing"]["empty"] is True
    assert out["best_chunks"]["best"] == [1, -1, 2]

    # rebuild_source: hide the second-best hit in the database, ask the program to rebuild its source
    disk = sqlite3.connect(path)
    disk.execute("UPDATE items SET hidden_at = 1 WHERE id = ?", (victim,))
    disk.commit()
    disk.close()
    out = _run(mirror_exe, "gpu", path, 7, qfile, live[victim][0])
    # build() itself now skips the hidden row, and so does the rebuilt source
    w_ids, w_scores = want(5, keep=ids != victim)
    g_ids, g_scores = _items(out["after_rebuild"])
    assert g_ids == w_ids and np.array_equal(g_scores, w_scores)


@pytest.mark.gpu
def test_mirror_searcher_over_two_gpus_in_one_process(mirror_exe, orc, tmp_path):
    """ONE perceive::Searcher over 2+ GPUs of one process (pcv_index_create_multi; the reference Searcher is
    one Send + Sync object, crates/perceive-tauri/src-tauri/app_state.rs:63-75): every step of the program
    prints exactly what the one-GPU Searcher prints, which test_mirror_searcher_matches_oracle holds to the oracle."""
    import torch
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip(f"needs 2 GPUs in one process, this box shows {n_gpu}")
    path, conn, live = _file_db(orc, tmp_path)
    q = orc.synth_rows(12, 0, 0, 1, DIM)[0]
    qfile = tmp_path / "query.f32"
    q.astype("<f4").tofile(qfile)
    some_source = live[sorted(live)[0]][0]
    one = _run(mirror_exe, "gpu", path, 7, qfile, some_source)
    devs = ",".join(str(d) for d in range(min(n_gpu, 4)))
    many = _run(mirror_exe, "gpu", path, 7, qfile, some_source, env={"PCV_MIRROR_DEVICES": devs})
    assert one["shards"]["world"] == 1 and many["shards"]["world"] == min(n_gpu, 4)
    assert many["shards"]["n_rows"] == one["shards"]["n_rows"] == len(live)
    for step in one:
        if step != "shards":
            assert many[step] == one[step], step
