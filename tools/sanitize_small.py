"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python tools/sanitize_small.py

No torch; sizes chosen so the run finishes in about a minute under the tool."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

import perceive_b200 as pb  # noqa: E402

os.environ.setdefault("PCV_GEMM_MIN_ROWS", "1")
rng = np.random.default_rng(0)


def unit(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def check(name, res, rows, qs, k, cosine=False, rtol=1e-5, atol=3e-6):
    truth = rows.astype(np.float64) @ qs.astype(np.float64).T
    if cosine:
        truth = truth / np.linalg.norm(rows.astype(np.float64), axis=1)[:, None] / np.linalg.norm(qs.astype(np.float64), axis=1)[None, :]
    ids, _, sims, cnt = res
    for b in range(qs.shape[0]):
        got = truth[ids[b, :cnt[b]] - 1, b]
        assert np.allclose(sims[b, :cnt[b]], got, rtol=rtol, atol=atol), name
        assert got.min() >= np.sort(truth[:, b])[-min(k, rows.shape[0])] - 1e-5, name
    print("ok", name, flush=True)


def bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7fff + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


n, d, k = 3000, 384, 10
rows, ids = unit(n, d), np.arange(1, n + 1, dtype=np.int64)
src = (np.arange(n) % 3).astype(np.int64)
with pb.Index(d) as ix:  # K1 fp32, source filter, small batch
    ix.set_rows(rows, ids, src)
    q = unit(5, d)
    check("K1 f32", ix.search(q, k), rows, q, k)
    ix.search(q[0], k, sources=[1])
    ix.replace_source(1, rows[:100], ids[:100] + 100000)
with pb.Index(d, store=pb.PCV_BF16, metric=pb.PCV_METRIC_COSINE) as ix:  # K1 bf16 cosine, K2 cosine
    r2 = rows * rng.uniform(0.5, 4.0, (n, 1)).astype(np.float32)
    ix.set_rows(r2, ids)
    q = unit(24, d)
    check("K1 bf16 cosine", ix.search(q[:2], k), bf16(r2), bf16(q[:2]), k, cosine=True)
    check("K2 bf16 cosine (pair)", ix.search(q, k), bf16(r2), bf16(q), k, cosine=True)
with pb.Index(d, store=pb.PCV_BF16) as ix:  # K2 dot, ragged tile counts, k = 100
    ix.set_rows(rows, ids, src)
    q = unit(130, d)
    check("K2 bf16", ix.search(q, 100), bf16(rows), bf16(q), 100)
    ix.search(q, 7, sources=[0, 2])
with pb.Index(768, store=pb.PCV_BF16) as ix:  # wide shape
    rw = unit(2000, 768)
    ix.set_rows(rw, np.arange(1, 2001, dtype=np.int64))
    q = unit(20, 768)
    check("K2 wide", ix.search(q, 50), bf16(rw), bf16(q), 50)
with pb.Index(d, store=pb.PCV_F32_SPLIT) as ix:  # K1 over two planes (grouped), K3 filter + rescoring, forced fallback
    ix.set_rows(rows, ids, src)
    q = unit(7, d)
    check("K1 split (grouped)", ix.search(q, k), rows, q, k)
    q = unit(140, d)
    check("K3 split batch", ix.search(q, k), rows, q, k)
    ix.search(q, 7, sources=[0, 2])
    same = np.tile(rows[:1], (n, 1))  # identical rows: the proof fails, every query takes the exact fallback scan
    ix.set_rows(same, ids)
    r = ix.search(q[:20], k)
    assert ix.stats().last_fallback_queries == 20 and np.array_equal(r[0][0], np.arange(1, k + 1))
    print("ok K3 fallback", flush=True)
    ix.replace_source(0, rows[:200], ids[:200] + 200000)
    ix.search(q[:20], k)
os.environ["PCV_GEMM_NO_PAIR"] = "1"
with pb.Index(d, store=pb.PCV_BF16) as ix:  # single-CTA kernel
    ix.set_rows(rows, ids)
    q = unit(40, d)
    check("K2 bf16 (single CTA)", ix.search(q, k), bf16(rows), bf16(q), k)
print("all ok")
