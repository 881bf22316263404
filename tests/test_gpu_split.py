"""Parity of K3 — batched fp32 search over PCV_F32_SPLIT rows (the fp32 values held exactly as a hi
and a lo 16-bit plane): tensor-core FILTER over the hi plane, exact fp32 rescoring of the candidates
in K1's summation order, proof of completeness, exact-scan fallback where the proof fails.

The bar is the integer one: ids, similarities and distances BIT-IDENTICAL to the fp32 scan (K1 on a
PCV_F32 index of the same rows) and to the oracle's restatement of that summation order
(oracle.c `dot_v1`), for friendly and adversarial inputs alike."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb(pcv_lib):
    import perceive_b200
    return perceive_b200


def _same_as_fp32_index(pb, rows, ids, qs, k, src=None, sources=None):
    """Search the same rows as a split index and as an fp32 index; everything must be bit-identical."""
    dim = rows.shape[1]
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(rows, ids, src)
        got = ix.search(qs, k, sources=sources)
        st = ix.stats()
    with pb.Index(dim, store=pb.PCV_F32) as ix:
        ix.set_rows(rows, ids, src)
        want = ix.search(qs, k, sources=sources)
    for g, w, name in zip(got, want, ("ids", "scores", "sims", "counts")):
        assert np.array_equal(g, w), f"split and fp32 index disagree on {name}"
    return got, st


@pytest.mark.parametrize("n,dim,nq,k", [(40_000, 384, 1, 10), (40_000, 384, 5, 10), (25_000, 384, 200, 10),
                                         (30_000, 128, 64, 100), (10_000, 100, 33, 7), (300, 384, 3, 50),
                                         (50_000, 768, 48, 10), (20_000, 384, 20, 200)])
def test_split_is_bit_identical_to_the_fp32_scan(pb, orc, n, dim, nq, k):
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(rows, ids)
        back, _, _ = ix.get_rows(0, min(n, 256))
    assert np.array_equal(back.view(np.uint32), rows[:back.shape[0]].view(np.uint32)), "the two planes must hold the fp32 value exactly"
    got, st = _same_as_fp32_index(pb, rows, ids, qs, k)
    filtered = nq >= 16 and n >= 4096 and k <= 128
    assert st.last_kernel == (2 if filtered else 1)
    for b in np.unique(np.linspace(0, nq - 1, 5).astype(int)):
        w_ids, w_scores, w_sims = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
        assert np.array_equal(got[0][b][:len(w_ids)], w_ids)
        assert np.array_equal(got[2][b][:len(w_ids)], w_sims.astype(np.float32))
        assert np.array_equal(got[1][b][:len(w_ids)], w_scores)
    print(f"split n={n} dim={dim} B={nq} k={k}: kernel {st.last_kernel}, {st.last_launches} launches, "
          f"{st.last_fallback_queries} queries re-done by the exact scan")


def test_split_filter_proves_most_queries_complete(pb, orc):
    """On unit-sphere data the proof should hold for (nearly) every query: the fallback is for adversarial inputs."""
    n, dim, nq, k = 200_000, 384, 256, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    got, st = _same_as_fp32_index(pb, rows, ids, qs, k)
    assert st.last_kernel == 2
    assert st.last_fallback_queries <= nq // 16, st.last_fallback_queries
    print(f"fallback queries: {st.last_fallback_queries} of {nq}")


def test_split_adversarial_inputs_stay_exact(pb, orc):
    """All-equal scores (every row identical: the filter's candidate set can never be proven complete, all
    queries go to the exact scan; ties resolve to the lowest ids), duplicated rows, and a corpus whose
    scores ascend with the row index (worst case for threshold pruning)."""
    dim, nq, k = 384, 32, 10
    rng = np.random.default_rng(5)
    qs = rng.standard_normal((nq, dim)).astype(np.float32)
    # 1. every row identical
    n = 20_000
    rows = np.tile(rng.standard_normal((1, dim)).astype(np.float32), (n, 1))
    ids = np.arange(1, n + 1, dtype=np.int64)
    got, st = _same_as_fp32_index(pb, rows, ids, qs, k)
    assert st.last_kernel == 2 and st.last_fallback_queries == nq
    assert np.array_equal(got[0], np.tile(np.arange(1, k + 1), (nq, 1)))
    # 2. each distinct row stored eight times (ties -> lower id)
    base = rng.standard_normal((3_000, dim)).astype(np.float32)
    rows = np.repeat(base, 8, axis=0)
    ids = np.arange(1, rows.shape[0] + 1, dtype=np.int64)
    got, st = _same_as_fp32_index(pb, rows, ids, qs, k)
    assert np.all(np.diff(got[0][:, :8], axis=1) == 1), "eight copies of the best row, lowest id first"
    # 3. scores ascending with the row index: row r = (r / n) * direction, queries = direction + noise
    n = 30_000
    d0 = rng.standard_normal(dim).astype(np.float32)
    d0 /= np.linalg.norm(d0)
    rows = (np.arange(1, n + 1, dtype=np.float32)[:, None] / n) * d0[None, :]
    qs3 = (d0[None, :] + 0.01 * rng.standard_normal((nq, dim))).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    got, st = _same_as_fp32_index(pb, rows, ids, qs3, k)
    w_ids, _, _ = orc.search(rows, ids, qs3[0], k, mode=orc.MODE_F32_V1)
    assert np.array_equal(got[0][0], w_ids)
    print(f"ascending corpus: {st.last_fallback_queries} of {nq} queries re-done by the exact scan")


def test_split_unnormalised_rows_and_sources(pb, orc):
    """Un-normalised rows (dot > dim clamps the reference distance to 0, order still by
    similarity) and a source filter, on split rows."""
    n, dim, nq, k = 20_000, 384, 40, 10
    rows = orc.synth_rows(1, orc.DIST_SCALED, 0, n, dim) * 8.0
    qs = orc.synth_rows(2, orc.DIST_SCALED, 0, nq, dim) * 8.0
    ids = np.arange(1, n + 1, dtype=np.int64)
    src = (np.arange(n) // 7000).astype(np.int64)
    for flt in (None, [0, 2], [1]):
        res, st = _same_as_fp32_index(pb, rows, ids, qs, k, src=src, sources=flt)
        sel = np.ones(n, bool) if flt is None else np.isin(src, flt)
        w_ids, _, _ = orc.search(rows[sel], ids[sel], qs[3], k, mode=orc.MODE_F32_V1)
        assert np.array_equal(res[0][3], w_ids)
    assert res[1].min() == 0.0  # clamp reached (search.rs:277)


def test_split_interleaved_ids_and_replace_source(pb, orc):
    """Two sources with interleaved ids (rank tables in play) and a segment swap on a two-plane matrix."""
    n, dim, nq, k = 12_000, 384, 24, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    src = (ids % 2).astype(np.int64)
    _same_as_fp32_index(pb, rows, ids, qs, k, src=src)
    new_rows = orc.synth_rows(7, 0, 0, 5_000, dim)
    new_ids = np.arange(100_001, 105_001, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(rows, ids, src)
        ix.replace_source(1, new_rows, new_ids)
        got = ix.search(qs, k)
        keep = src == 0
        back, bids, bsrc = ix.get_rows(0, int(keep.sum()) + 5_000)
    all_rows = np.concatenate([rows[keep], new_rows])
    all_ids = np.concatenate([ids[keep], new_ids])
    assert np.array_equal(back.view(np.uint32), all_rows.view(np.uint32)) and np.array_equal(bids, all_ids)
    for b in (0, nq - 1):
        w_ids, _, w_sims = orc.search(all_rows, all_ids, qs[b], k, mode=orc.MODE_F32_V1)
        assert np.array_equal(got[0][b], w_ids) and np.array_equal(got[2][b], w_sims.astype(np.float32))


def test_split_batch_larger_than_one_chunk(pb, orc):
    """More than 4096 queries: the filter runs in chunks, rescoring and the fallback list keep GLOBAL query indices."""
    n, dim, nq, k = 9_000, 128, 4_200, 5
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    for b in (0, 1, 4095, 4096, 4097, nq - 1):
        w_ids, w_scores, w_sims = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
        assert np.array_equal(res[0][b], w_ids) and np.array_equal(res[2][b], w_sims.astype(np.float32)), b
    print(f"chunked split batch: {st.last_launches} launches, {st.last_fallback_queries} fallback queries")


def test_split_rejects_what_it_does_not_implement(pb):
    with pytest.raises(pb.PcvError) as e:
        pb.Index(384, store=pb.PCV_F32_SPLIT, metric=pb.PCV_METRIC_COSINE)
    assert e.value.code == 5
    with pytest.raises(pb.PcvError):
        pb.Index(1024, store=pb.PCV_F32_SPLIT)
    with pb.Index(384, store=pb.PCV_F32_SPLIT) as ix:
        big = np.eye(384, dtype=np.float32)
        big[0, 0] = 3.4e38  # finite and huge: the hi plane truncates, never rounds up to infinity — legal
        ix.set_rows(big, np.arange(384))
        back, _, _ = ix.get_rows(0, 1)
        assert back[0, 0] == np.float32(3.4e38)
        big[1, 1] = np.inf
        with pytest.raises(pb.PcvError) as e:
            ix.set_rows(big, np.arange(384))
        assert e.value.code == 3


def test_split_config4_shape_subsample(pb, orc):
    """BASELINE config 4's shape (batch 256, 384-d, top-10) on 300k device-generated rows: every
    sampled query identical, bit for bit, to the oracle's fp32 scan."""
    n, dim, nq, k = 300_000, 384, 256, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.generate_synthetic(n, seed=1)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    for b in range(0, nq, 16):
        w_ids, w_scores, w_sims = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
        assert np.array_equal(res[0][b], w_ids) and np.array_equal(res[2][b], w_sims.astype(np.float32))
        assert np.array_equal(res[1][b], w_scores)
    print(f"K3 config-4 shape: 16/16 sampled queries bit-identical to the fp32 scan; {st.last_launches} launches, "
          f"{st.last_search_ms:.3f} ms, {st.last_fallback_queries} fallback queries")
