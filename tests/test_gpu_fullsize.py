"""BASELINE's full sizes, checked through size-independent properties (the float64 oracle cannot
scan 10^7-10^8 rows x 10^3 queries in seconds):

  * planted queries: query j is a stored row p(j) fetched back from the index, so the exact top-1
    is known a priori (the row itself, similarity = |row|^2, cosine = 1) whatever the corpus size;
  * the k results are sorted by the stated order and every returned similarity is reproduced by a
    float64 dot against the returned row (spot check of the whole pipeline: ids, keys, emit);
  * batch-split invariance: the same queries searched as one batch and as two halves give the same ids.
Corpora are generated on the device (pcv_index_generate_synthetic); rows are regenerated on the host
by the oracle for the spot checks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb(pcv_lib):
    import perceive_b200
    return perceive_b200


def _planted(ix, n, nq):
    pos = (np.arange(nq, dtype=np.int64) * 7919 + 13) % n  # distinct for nq < n
    qs = np.concatenate([ix.get_rows(int(p), 1)[0] for p in pos])
    return pos, qs


def _spot_check(orc, res, qs, seed, dist, dim, round_bf16, cosine, k, n_spot=6):
    ids, scores, sims, cnt = res
    assert np.all(cnt == k)
    assert np.all(np.diff(sims, axis=1) <= 0), "similarities must be descending"
    for b in np.linspace(0, qs.shape[0] - 1, n_spot).astype(int):
        rows = np.concatenate([orc.synth_rows(seed, dist, int(i) - 1, 1, dim) for i in ids[b]])
        q = qs[b:b + 1]
        if round_bf16:
            rows, q = orc.round_bf16(rows), orc.round_bf16(q)
        t = rows.astype(np.float64) @ q[0].astype(np.float64)
        if cosine:
            t = t / np.linalg.norm(rows.astype(np.float64), axis=1) / np.linalg.norm(q[0].astype(np.float64))
        np.testing.assert_allclose(sims[b], t, rtol=1e-5, atol=2e-6)


def test_config2_full_size_planted(pb, orc):
    """1M x 384 fp32, top-10 (K1): planted row first with distance exactly 1 - |row|^2/384."""
    n, dim, k = 1_000_000, 384, 10
    with pb.Index(dim) as ix:
        ix.generate_synthetic(n, seed=1)
        pos, qs = _planted(ix, n, 8)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 1
    assert np.array_equal(res[0][:, 0], pos + 1)
    _spot_check(orc, res, qs, 1, 0, dim, False, False, k)


def test_config3_full_size_planted(pb, orc):
    """10M x 384 bf16, batch 1024, top-100 (K2 pair kernel: bootstrap pass, then three threshold passes)."""
    n, dim, nq, k = 10_000_000, 384, 1024, 100
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.generate_synthetic(n, seed=1)
        pos, qs = _planted(ix, n, nq)
        res = ix.search(qs, k)
        st = ix.stats()
        assert st.last_kernel == 2 and st.last_launches >= 8, st.last_launches
        halves = [ix.search(qs[:512], k), ix.search(qs[512:], k)]
    assert np.array_equal(res[0][:, 0], pos + 1), "a stored row must be its own nearest neighbour"
    assert np.allclose(res[2][:, 0], 1.0, atol=2e-2)  # |bf16(unit row)|^2
    assert np.array_equal(np.concatenate([h[0] for h in halves]), res[0]), "batch-split invariance"
    _spot_check(orc, res, qs, 1, 0, dim, True, False, k)


def test_config4_shard_size_planted(pb, orc):
    """25M x 384 fp32 rows held as two 16-bit planes (config 4's per-GPU shard at 4 GPUs), batch 256, top-10 (K3:
    tensor-core filter + exact rescoring); a planted row's similarity with itself is its exact fp32 norm."""
    n, dim, nq, k = 25_000_000, 384, 256, 10
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.generate_synthetic(n, seed=1)
        pos, qs = _planted(ix, n, nq)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
    assert np.array_equal(res[0][:, 0], pos + 1)
    assert np.allclose(res[2][:, 0], 1.0, atol=1e-6)
    _spot_check(orc, res, qs, 1, 0, dim, False, False, k)
    # exact, not just close: the oracle's fp32 scan order on the returned rows reproduces every bit
    for b in (0, nq // 2, nq - 1):
        rows_b = np.concatenate([orc.synth_rows(1, 0, int(i) - 1, 1, dim) for i in res[0][b]])
        want = np.array([orc.dot(qs[b], r, mode=orc.MODE_F32_V1) for r in rows_b], dtype=np.float32)
        assert np.array_equal(res[2][b], want)


def test_config5_shard_size_planted(pb, orc):
    """6.25M x 768 bf16 un-normalised rows (config 5's per-GPU shard), batch 4096, top-50, cosine."""
    n, dim, nq, k = 6_250_000, 768, 4096, 50
    with pb.Index(dim, store=pb.PCV_BF16, metric=pb.PCV_METRIC_COSINE) as ix:
        ix.generate_synthetic(n, seed=1, dist=pb.PCV_DIST_SCALED)
        pos, qs = _planted(ix, n, nq)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
    assert np.array_equal(res[0][:, 0], pos + 1)
    assert np.allclose(res[2][:, 0], 1.0, atol=1e-5), "cosine of a row with itself"
    _spot_check(orc, res, qs, 1, 1, dim, True, True, k)
