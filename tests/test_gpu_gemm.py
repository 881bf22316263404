"""Parity of K2 — the tcgen05/TMEM batched search — against the float64 oracle.

K2 runs when a bf16 index is searched with a batch of >= 16 queries.  Both
operands are bf16 (the index stores bf16 rows; queries are rounded to bf16 on
entry), every product is exact in fp32, and the tensor core accumulates in fp32
in an order the hardware does not document.  The bar is therefore:
  * similarities within GEMM_RTOL * |sim| + GEMM_ATOL of the float64 dot of the
    same bf16 values (stated tolerance, measured max error is printed);
  * the returned set is the exact top-k up to epsilon-ties: every returned row
    scores >= (k-th true score - eps) and every row scoring > (k-th + eps) is
    returned (recall@k = 1.0 against the float64 ranking outside the tie band);
  * order: similarity descending, exact ties by lower id.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GEMM_RTOL = 1e-5
GEMM_ATOL = 2e-6


@pytest.fixture(scope="module")
def pb(pcv_lib):
    import perceive_b200
    return perceive_b200


def check_batch(res, stored, ids, queries_bf16, k, *, selected=None, what="", rtol=None, atol=None, cosine=False):
    """res = Index.search output; stored = bf16-rounded rows (fp32 array);
    queries_bf16 = bf16-rounded queries; selected = boolean row mask (source filter)."""
    GEMM_RTOL = globals()["GEMM_RTOL"] if rtol is None else rtol
    GEMM_ATOL = globals()["GEMM_ATOL"] if atol is None else atol
    g_ids, g_scores, g_sims, g_cnt = res
    rows64 = stored.astype(np.float64)
    n = rows64.shape[0]
    sel = np.ones(n, dtype=bool) if selected is None else selected
    sel_idx = np.nonzero(sel)[0]
    q64 = queries_bf16.astype(np.float64)
    truth = rows64[sel_idx] @ q64.T  # [n_sel, B]
    if cosine:  # crates/perceive-core/lib.rs:67-77: rows and queries divided by their L2 norms
        truth = truth / np.linalg.norm(rows64[sel_idx], axis=1)[:, None] / np.linalg.norm(q64, axis=1)[None, :]
    id_of = ids[sel_idx]
    pos_of_id = {int(i): j for j, i in enumerate(id_of)}
    max_err = 0.0
    dim = stored.shape[1]
    for b in range(queries_bf16.shape[0]):
        t = truth[:, b]
        want_cnt = min(k, len(sel_idx))
        assert int(g_cnt[b]) == want_cnt, f"{what} q{b}: count {g_cnt[b]} != {want_cnt}"
        c = want_cnt
        gi, gs = g_ids[b, :c], g_sims[b, :c].astype(np.float64)
        assert len(set(gi.tolist())) == c, f"{what} q{b}: duplicate ids"
        assert np.all(g_ids[b, c:] == -1) and np.all(np.isinf(g_scores[b, c:]))
        tp = np.array([t[pos_of_id[int(i)]] for i in gi])
        err = np.abs(gs - tp)
        tol = GEMM_RTOL * np.abs(tp) + GEMM_ATOL
        max_err = max(max_err, float(err.max()) if c else 0.0)
        assert np.all(err <= tol), f"{what} q{b}: similarity error {err.max():.3e} above tolerance"
        if c == 0:
            continue
        order = np.lexsort((id_of, -t))
        kth = t[order[c - 1]]
        eps = 4 * (GEMM_RTOL * abs(kth) + GEMM_ATOL)
        assert np.all(tp >= kth - eps), f"{what} q{b}: returned a row below the k-th true score"
        must = set(id_of[t > kth + eps].tolist())
        assert must <= set(gi.tolist()), f"{what} q{b}: missed {len(must - set(gi.tolist()))} rows above the tie band"
        # order: descending by the GPU's own similarity; exact ties by lower id
        d = np.diff(g_sims[b, :c])
        assert np.all(d <= 0), f"{what} q{b}: similarities not descending"
        tie = np.nonzero(d == 0)[0]
        assert np.all(gi[tie] < gi[tie + 1]), f"{what} q{b}: tie not broken by lower id"
        # reported score = reference distance of the GPU similarity (search.rs:274-277)
        if cosine:  # reported score = the similarity itself
            assert np.array_equal(g_scores[b, :c], g_sims[b, :c])
        else:
            want_sc = np.maximum(np.float32(1.0) - g_sims[b, :c] / np.float32(dim), np.float32(0.0))
            assert np.array_equal(g_scores[b, :c], want_sc.astype(np.float32))
    return max_err


def _make(orc, n, dim, nq, dist=0):
    rows = orc.synth_rows(1, dist, 0, n, dim)
    stored = orc.round_bf16(rows)
    qs = orc.round_bf16(orc.synth_rows(2, dist, 0, nq, dim))
    ids = np.arange(1, n + 1, dtype=np.int64)
    return rows, stored, qs, ids


@pytest.mark.parametrize("n,dim,nq,k", [(50_000, 384, 64, 10), (50_000, 384, 130, 100), (30_000, 128, 16, 32),
                                         (20_000, 100, 40, 7), (10_000, 64, 128, 128), (33_333, 256, 300, 50)])
def test_gemm_batch_matches_float64_truth(pb, orc, n, dim, nq, k):
    rows, stored, qs, ids = _make(orc, n, dim, nq)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2, "the tcgen05 path did not run"
    err = check_batch(res, stored, ids, qs, k, what=f"n={n} dim={dim} B={nq} k={k}")
    print(f"K2 max |sim - f64| = {err:.3e}")


def test_gemm_multi_pass_schedule(pb, orc, monkeypatch):
    """Small candidate buffers + pass ratio 2 force several threshold passes."""
    monkeypatch.setenv("PCV_GEMM_CAND_CAP", "512")
    monkeypatch.setenv("PCV_GEMM_PASS_RATIO", "2")
    monkeypatch.setenv("PCV_GEMM_BOOT_TILES", "0")  # the plain geometric schedule is the subject here
    n, dim, nq, k = 150_000, 384, 200, 100
    rows, stored, qs, ids = _make(orc, n, dim, nq)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2 and st.last_launches >= 7, st.last_launches  # >= 3 passes x (gemm + select) + convert
    check_batch(res, stored, ids, qs, k, what="multi-pass")


def test_gemm_overflow_prune_path_adversarial(pb, orc, monkeypatch):
    """Scores increasing with the row index: every score beats the running
    threshold, so candidate buffers overflow and are pruned in-kernel; results stay exact."""
    monkeypatch.setenv("PCV_GEMM_MAX_CTAS", "3")
    monkeypatch.setenv("PCV_GEMM_CAND_CAP", "512")
    n, dim, nq, k = 40_000, 128, 48, 20
    v = orc.round_bf16(orc.synth_rows(3, 0, 0, 1, dim))[0]
    scale = orc.round_bf16((1.0 + np.arange(n, dtype=np.float32) / 64.0).reshape(n, 1))
    rows = (scale * v[None, :]).astype(np.float32)
    stored = orc.round_bf16(rows)
    qs = orc.round_bf16(np.abs(orc.synth_rows(2, 0, 0, nq, dim)) * np.sign(v)[None, :])  # every q.v > 0
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    check_batch(res, stored, ids, qs, k, what="ascending")


def test_gemm_sources_filter_and_interleaved_ids(pb, orc):
    """Three sources with interleaved ids: the filter becomes row ranges, ids go
    through the rank tables (tie-break by id, not by row)."""
    n, dim, nq, k = 24_000, 192, 32, 25
    rows, stored, qs, _ = _make(orc, n, dim, nq)
    rng = np.random.default_rng(5)
    ids = rng.permutation(np.arange(10, 10 + n)).astype(np.int64)
    src = (np.arange(n) % 3).astype(np.int64) * 7  # sources 0, 7, 14 interleaved
    rows[100] = rows[5]  # exact duplicates across sources: ties by id
    rows[101] = rows[5]
    stored = orc.round_bf16(rows)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids, src)
        for flt in ([0, 14], [7], [0, 7, 14], [99]):
            res = ix.search(qs, k, sources=flt)
            st = ix.stats()
            mask = np.isin(src, flt)
            if mask.sum() >= 4096:
                assert st.last_kernel == 2
            check_batch(res, stored, ids, qs, k, selected=mask, what=f"sources={flt}")


def test_gemm_fewer_rows_than_k(pb, orc, monkeypatch):
    monkeypatch.setenv("PCV_GEMM_MIN_ROWS", "1")
    n, dim, nq, k = 300, 384, 20, 100
    rows, stored, qs, ids = _make(orc, n, dim, nq)
    src = (np.arange(n) >= 250).astype(np.int64)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids, src)
        res = ix.search(qs, k, sources=[1])  # 50 rows < k
        assert ix.stats().last_kernel == 2
        check_batch(res, stored, ids, qs, k, selected=src == 1, what="rows<k")
        res = ix.search(qs, k, sources=[5])  # nothing selected
        assert np.all(res[3] == 0) and np.all(res[0] == -1)


def test_gemm_agrees_with_scan_on_the_same_bf16_index(pb, orc):
    """K1 (one query at a time) and K2 (batched) score the same bf16 operands:
    same ids outside epsilon-ties, similarities within the stated tolerance."""
    n, dim, nq, k = 60_000, 384, 24, 10
    rows, stored, qs, ids = _make(orc, n, dim, nq)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        batched = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
        for b in range(nq):
            one = ix.search(qs[b], k)
            assert ix.stats().last_kernel == 1
            np.testing.assert_allclose(one[2][0], batched[2][b], rtol=GEMM_RTOL, atol=GEMM_ATOL)
            if not np.array_equal(one[0][0], batched[0][b]):
                gaps = np.abs(np.diff(one[2][0]))
                assert gaps.min() <= 4 * GEMM_ATOL, f"query {b}: ids differ without an epsilon-tie"


def test_gemm_config3_shape_subsample(pb, orc):
    """BASELINE config 3's shape (batch 1024, 384-d bf16, top-100) on a corpus the
    float64 truth finishes in seconds (200k rows), generated on the device."""
    n, dim, nq, k = 200_000, 384, 1024, 100
    stored = orc.round_bf16(orc.synth_rows(1, 0, 0, n, dim))
    qs = orc.round_bf16(orc.synth_rows(2, 0, 0, nq, dim))
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.generate_synthetic(n, seed=1)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    err = check_batch(res, stored, ids, qs, k, what="config3-shape")
    print(f"K2 config-3 shape: max |sim - f64| = {err:.3e}, {st.last_launches} launches, {st.last_search_ms:.3f} ms")


@pytest.mark.parametrize("n,dim,nq,k", [(30_000, 768, 64, 50), (20_000, 512, 130, 10), (9_000, 440, 20, 100)])
def test_gemm_wide_rows_up_to_768(pb, orc, n, dim, nq, k):
    """384 < dim <= 768 (distilbert-shaped, BASELINE config 5): 12 K blocks, 64-row document tiles."""
    rows, stored, qs, ids = _make(orc, n, dim, nq)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
    err = check_batch(res, stored, ids, qs, k, what=f"wide dim={dim}")
    print(f"K2 wide dim={dim}: max |sim - f64| = {err:.3e}")


@pytest.mark.parametrize("n,dim,nq,k", [(30_000, 768, 96, 50), (30_000, 384, 40, 10)])
def test_gemm_cosine_unnormalised_rows(pb, orc, n, dim, nq, k):
    """BASELINE config 5's semantics: un-normalised bf16 rows, cosine (lib.rs:67-77) with the row norms
    computed on the device from the stored values and applied in the epilogue."""
    rows, stored, qs, ids = _make(orc, n, dim, nq, dist=1)  # DIST_SCALED: per-row scales in [0.25, 8)
    with pb.Index(dim, store=pb.PCV_BF16, metric=pb.PCV_METRIC_COSINE) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
        one = ix.search(qs[0], k)  # the scan (K1) computes norms in-kernel: same cosine within tolerance
        assert ix.stats().last_kernel == 1
    err = check_batch(res, stored, ids, qs, k, what=f"cosine dim={dim}", cosine=True)
    np.testing.assert_allclose(one[2][0], res[2][0], rtol=GEMM_RTOL, atol=GEMM_ATOL)
    print(f"K2 cosine dim={dim}: max |cos - f64| = {err:.3e}")


def test_gemm_batch_larger_than_one_chunk(pb, orc):
    """More than 4096 queries run as consecutive chunks; every query still gets its exact top-k."""
    n, dim, nq, k = 6_000, 64, 4_300, 5
    rows, stored, qs, ids = _make(orc, n, dim, nq)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
    check_batch(res, stored, ids, qs, k, what="chunked batch")


def test_gemm_config5_shape_subsample(pb, orc):
    """BASELINE config 5's shape (batch 4096, 768-d bf16, top-50, un-normalised rows, cosine) on 12k rows."""
    n, dim, nq, k = 12_000, 768, 4_096, 50
    rows, stored, qs, ids = _make(orc, n, dim, nq, dist=1)
    with pb.Index(dim, store=pb.PCV_BF16, metric=pb.PCV_METRIC_COSINE) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    err = check_batch(res, stored, ids, qs, k, what="config5-shape", cosine=True)
    print(f"K2 config-5 shape: max |cos - f64| = {err:.3e}, {st.last_launches} launches, {st.last_search_ms:.3f} ms")


@pytest.mark.parametrize("store_name,dim,nq,k", [("bf16", 384, 70, 10), ("bf16", 768, 20, 50), ("split", 384, 40, 10)])
def test_single_cta_kernel_matches_too(pb, orc, monkeypatch, store_name, dim, nq, k):
    """PCV_GEMM_NO_PAIR forces the one-CTA tcgen05 kernel (normally only used when a pass has a
    single tile): same results, same tolerance."""
    monkeypatch.setenv("PCV_GEMM_NO_PAIR", "1")
    n = 9_000
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    if store_name == "bf16":
        store, rows_ref, qs_ref = pb.PCV_BF16, orc.round_bf16(rows), orc.round_bf16(qs)
    else:
        store, rows_ref, qs_ref = pb.PCV_F32_SPLIT, rows, qs
    with pb.Index(dim, store=store) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 2
    check_batch(res, rows_ref, ids, qs_ref, k, what=f"single-CTA {store_name} dim={dim}")


@pytest.mark.parametrize("store_name", ["bf16", "split"])
def test_gemm_hidden_rows_are_cut_out(pb, orc, store_name):
    """pcv_index_set_hidden on the tensor path: hidden rows split the row ranges, so tiles start and end
    at arbitrary rows; results equal the truth over the visible rows only."""
    n, dim, nq, k = 24_000, 384, 32, 20
    rows, stored, qs, _ = _make(orc, n, dim, nq)
    rng = np.random.default_rng(3)
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64)
    src = (np.arange(n) % 2).astype(np.int64)
    split = store_name == "split"
    store = pb.PCV_F32_SPLIT if split else pb.PCV_BF16
    if split:
        stored, qq = rows, orc.synth_rows(2, 0, 0, nq, dim)
        tol = dict(rtol=1e-5, atol=2e-6)
    else:
        qq, tol = qs, {}
    with pb.Index(dim, store=store) as ix:
        ix.set_rows(rows, ids, src)
        first = ix.search(qq, k)
        hide = set(int(i) for i in first[0][:4, :5].ravel())        # best hits of four queries
        hide |= set(int(i) for i in rng.choice(ids, 200, replace=False))
        ix.set_hidden(sorted(hide))
        mask = ~np.isin(ids, list(hide))
        for flt in (None, [1]):
            res = ix.search(qq, k, sources=flt)
            assert ix.stats().last_kernel == 2
            sel = mask if flt is None else mask & np.isin(src, flt)
            check_batch(res, stored, ids, qq, k, selected=sel, what=f"{store_name} hidden sources={flt}", **tol)
            assert not set(res[0].ravel().tolist()) & hide


@pytest.mark.parametrize("store_name,n,dim,nq,k,cosine", [("bf16", 100_000, 384, 64, 10, False), ("bf16", 140_000, 384, 130, 30, False),
                                                           ("split", 120_000, 384, 40, 10, False), ("bf16", 80_000, 768, 48, 12, True)])
def test_gemm_bootstrap_pass(pb, orc, monkeypatch, store_name, n, dim, nq, k, cosine):
    """Large corpora open with a bootstrap pass (tile maxima only -> k-th largest = first threshold).  A small
    PCV_GEMM_BOOT_TILES makes these corpora take it; results must equal the truth and the plain schedule's."""
    dist = 1 if cosine else 0
    rows, stored, qs, ids = _make(orc, n, dim, nq, dist=dist)
    split = store_name == "split"
    if split:
        stored, qs = rows, orc.synth_rows(2, 0, 0, nq, dim)
    tol = dict(rtol=1e-5, atol=2e-6) if split else {}
    metric = pb.PCV_METRIC_COSINE if cosine else pb.PCV_METRIC_DOT_REF
    with pb.Index(dim, store=pb.PCV_F32_SPLIT if split else pb.PCV_BF16, metric=metric) as ix:
        ix.set_rows(rows, ids)
        monkeypatch.setenv("PCV_GEMM_BOOT_TILES", "0")  # off: the geometric schedule from pass 0
        plain = ix.search(qs, k)
        plain_launches = ix.stats().last_launches
        kk = max(48, 2 * k + 28) if split else k  # a split search filters for more candidates than it returns
        monkeypatch.setenv("PCV_GEMM_BOOT_TILES", str(max(4 * kk, 64)))
        res = ix.search(qs, k)
        st = ix.stats()
    # sizes chosen so that the plain schedule needs four passes and the bootstrapped one two (+ the 2 bootstrap
    # launches); the first cosine search also computes the row norms once
    assert st.last_kernel == 2 and st.last_launches < plain_launches, (st.last_launches, plain_launches)
    check_batch(res, stored, ids, qs, k, what=f"bootstrap {store_name} n={n} k={k}", cosine=cosine, **tol)
    assert np.array_equal(res[0], plain[0]) and np.array_equal(res[2], plain[2]) and np.array_equal(res[3], plain[3])
