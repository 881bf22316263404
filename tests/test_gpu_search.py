"""Parity of the CUDA path (through the C ABI) against the oracle — GPU only."""
import numpy as np
import pytest

from helpers import assert_same_result

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb(pcv_lib):
    import perceive_b200
    return perceive_b200


def _one(res, b=0):
    ids, scores, sims, counts = res
    return ids[b], scores[b], sims[b], counts[b]


@pytest.mark.parametrize("n,dim,k", [(10_000, 384, 10), (1, 384, 10), (7, 384, 10), (1000, 768, 20),
                                      (5000, 100, 5), (3000, 1024, 33), (2000, 16, 1), (20_000, 384, 100),
                                      (4000, 384, 128), (3000, 384, 300)])
def test_scan_f32_matches_oracle_bitexact(pb, orc, n, dim, k):
    """BASELINE config 1 shape (10k x 384, top-10) and ragged variants: ids, fp32
    similarities and reference distances identical to the oracle's v1-order scan."""
    rows = orc.synth_rows(1, 0, 0, n, dim)
    q = orc.synth_rows(2, 0, 0, 1, dim)[0]
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, k))
    want = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1, epc=4)
    assert_same_result(got, want, what=f"n={n} dim={dim} k={k}")
    # and within 1e-5 relative of the float64 truth (north_star tolerance)
    truth = orc.np_search(rows, ids, q, k)
    c = int(got[3])
    np.testing.assert_allclose(got[2][:c], truth[2][:c], rtol=1e-5, atol=1e-7)


def test_config1_matches_the_committed_golden_fixture(pb, orc):
    """BASELINE configs[0] through the C ABI against tests/golden/config1_top10.json (made by
    tests/golden/make_golden.py from the oracle): ids, similarities and reference distances, bit for bit."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "config1_top10.json").read_text())
    n, dim, k = g["rows"], g["dim"], g["k"]
    rows = orc.synth_rows(g["corpus_seed"], 0, 0, n, dim)
    q = orc.synth_rows(g["query_seed"], 0, 0, 1, dim)[0]
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got_ids, got_scores, got_sims, got_cnt = _one(ix.search(q, k))
    assert int(got_cnt) == k
    assert got_ids.tolist() == g["ids"]
    assert np.ascontiguousarray(got_sims, dtype=np.float32).view(np.uint32).tolist() == g["sim_bits"]
    assert np.ascontiguousarray(got_scores, dtype=np.float32).view(np.uint32).tolist() == g["score_bits"]


def test_synthetic_generator_matches_oracle(pb, orc):
    """Device generator == host generator == oracle restatement, bit for bit."""
    n, dim = 513, 384
    with pb.Index(dim) as ix:
        ix.generate_synthetic(n, seed=1, first_row=100)
        rows, ids, src = ix.get_rows(0, n)
    want = orc.synth_rows(1, 0, 100, n, dim)
    assert np.array_equal(rows, want)
    assert np.array_equal(ids, np.arange(101, 101 + n))
    host = np.empty((4, dim), dtype=np.float32)
    from perceive_b200 import _ffi
    _ffi.check(_ffi.load().pcv_synthetic_rows_host(1, 0, 100, 4, dim, host.ctypes.data))
    assert np.array_equal(host, want[:4])


def test_config2_shape_1m_rows(pb, orc):
    """BASELINE config 2: 1 query vs 1M x 384 fp32, top-10, generated on the device;
    checked against the oracle's scan of the identical host-generated corpus."""
    n, dim, k = 1_000_000, 384, 10
    q = orc.synth_rows(2, 0, 0, 1, dim)[0]
    with pb.Index(dim) as ix:
        ix.generate_synthetic(n, seed=1)
        got = _one(ix.search(q, k))
        st = ix.stats()
    assert st.last_kernel == 1 and st.last_launches == 1
    assert st.last_scan_bytes == n * dim * 4
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    want = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1)
    assert_same_result(got, want, what="C2")
    # the timed CPU baseline (baseline.c) agrees on the ids and within 1e-5 on the sims
    f_ids, f_scores, f_sims = orc.search_fast(rows, q, k)
    assert np.array_equal(f_ids, want[0])
    np.testing.assert_allclose(f_sims, want[2], rtol=1e-5, atol=1e-7)


def test_ties_break_to_lower_id(pb, orc):
    """Duplicated rows: equal similarities resolve to the lower doc id first."""
    dim = 384
    base = orc.synth_rows(3, 0, 0, 50, dim)
    rows = np.concatenate([base, base, base])  # every row three times
    ids = np.concatenate([np.arange(1000, 1050), np.arange(10, 60), np.arange(500, 550)]).astype(np.int64)
    q = base[7]
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, 6))
    want = orc.search(rows, ids, q, 6, mode=orc.MODE_F32_V1)
    assert_same_result(got, want, what="ties")
    assert list(got[0][:3]) == [17, 507, 1007]


def test_all_equal_scores(pb, orc):
    dim, n = 384, 3000
    row = orc.synth_rows(4, 0, 0, 1, dim)
    rows = np.repeat(row, n, axis=0)
    ids = np.random.default_rng(0).permutation(np.arange(1, n + 1)).astype(np.int64)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(row[0], 10))
    assert list(got[0]) == list(range(1, 11))


def test_ascending_scores_adversarial(pb, orc):
    """Similarity increases with the row index: every row beats the running
    threshold, the worst case for the fused top-k."""
    dim, n, k = 384, 20_000, 10
    q = orc.synth_rows(2, 0, 0, 1, dim)[0]
    scale = np.linspace(0.1, 1.0, n, dtype=np.float32)[:, None]
    rows = (q[None, :] * scale).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, k))
    want = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1)
    assert_same_result(got, want, what="ascending")


def test_sources_filter_and_interleaved_ids(pb, orc):
    """Per-source segments with interleaved ids (search.rs:166 filter); duplicates
    across sources must still tie-break on the doc id, not on storage order."""
    dim = 384
    rng = np.random.default_rng(5)
    n = 6000
    rows = orc.synth_rows(6, 0, 0, n, dim)
    rows[3000:3100] = rows[100:200]  # same documents present in two sources
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64)
    src = rng.integers(0, 4, size=n).astype(np.int64) * 10 + 3  # sources 3,13,23,33
    src[100:200] = 33
    src[3000:3100] = 3
    q = rows[150]
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids, src)
        for flt in (None, [3], [13, 33], [3, 13, 23, 33], [99], []):
            got = _one(ix.search(q, 12, sources=flt))
            want = orc.search(rows, ids, q, 12, source_ids=src, sources=flt, mode=orc.MODE_F32_V1)
            assert_same_result(got, want, what=f"filter={flt}")
            truth = orc.np_search(rows, ids, q, 12, source_ids=src, sources=flt)
            assert np.array_equal(got[0][: int(got[3])], truth[0])


def test_k_larger_than_rows(pb, orc):
    dim = 384
    rows = orc.synth_rows(7, 0, 0, 5, dim)
    ids = np.array([50, 40, 30, 20, 10], dtype=np.int64)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(rows[0], 10))
    want = orc.search(rows, ids, rows[0], 10, mode=orc.MODE_F32_V1)
    assert int(got[3]) == 5
    assert_same_result(got, want, what="k>N")


def test_empty_index_and_unknown_source(pb):
    with pb.Index(384) as ix:
        ix.set_rows(np.zeros((0, 384), np.float32), np.zeros(0, np.int64))
        ids, scores, sims, counts = ix.search(np.ones(384, np.float32), 10)
    assert counts[0] == 0 and np.all(ids == -1)


def test_identity_corpus(pb, orc):
    """rows = +-e_i: the answer is known a priori."""
    dim = 384
    rows = np.concatenate([np.eye(dim, dtype=np.float32), -np.eye(dim, dtype=np.float32)])
    ids = np.arange(1, 2 * dim + 1, dtype=np.int64)
    q = np.zeros(dim, np.float32)
    q[5], q[9], q[300] = 0.9, -0.8, 0.7
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, 3))
    assert list(got[0]) == [6, dim + 10, 301]
    np.testing.assert_array_equal(got[2], np.array([0.9, 0.8, 0.7], np.float32))
    np.testing.assert_array_equal(got[1], orc.np_distance(np.array([0.9, 0.8, 0.7], np.float32), dim))


def test_unnormalised_dot_clamps_distance(pb, orc):
    """dot > dim: the reference distance clamps to 0 (search.rs:277) but the order
    is still by similarity."""
    dim = 16
    rows = np.stack([np.full(dim, v, np.float32) for v in (3.0, 2.0, 1.5, 0.1)])
    ids = np.array([4, 3, 2, 1], dtype=np.int64)
    q = np.ones(dim, np.float32)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, 4))
    assert list(got[0]) == [4, 3, 2, 1]
    assert list(got[1][:3]) == [0.0, 0.0, 0.0] and got[1][3] > 0.8


def test_nonfinite_inputs_rejected(pb):
    rows = np.ones((4, 384), np.float32)
    rows[2, 7] = np.nan
    with pb.Index(384) as ix:
        with pytest.raises(pb.PcvError) as e:
            ix.set_rows(rows, np.arange(4))
        assert e.value.code == 3
        ix.set_rows(np.ones((4, 384), np.float32), np.arange(4))
        q = np.ones(384, np.float32)
        q[0] = np.inf
        with pytest.raises(pb.PcvError) as e:
            ix.search(q, 2)
        assert e.value.code == 3


def test_small_batch_matches_single_queries(pb, orc):
    """Batched entry point == one search_vector per query (bit-identical)."""
    n, dim, k = 30_000, 384, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = orc.synth_rows(2, 0, 0, 7, dim)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        for b in range(7):
            want = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
            assert_same_result(_one(res, b), want, what=f"batch query {b}")


@pytest.mark.parametrize("nq,k", [(9, 100), (270, 10), (70, 40)])
def test_grouped_scan_walks_batches_over_fp32_rows(pb, orc, nq, k):
    """Batches over fp32 rows run as GROUPED launches (a list of queries, four per pass, up to 64 groups = 256
    queries per launch): every query still gets exactly what the single-query scan returns, in 1 / 2 / 1 launches."""
    n, dim = 12_000, 384
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        st = ix.stats()
        assert st.last_kernel == 1 and st.last_launches == (nq + 255) // 256, st.last_launches
        for b in np.unique(np.linspace(0, nq - 1, 6).astype(int)):
            want = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
            assert_same_result(_one(res, b), want, what=f"grouped batch query {b}")
            single = ix.search(qs[b], k)
            assert np.array_equal(single[0][0], res[0][b]) and np.array_equal(single[2][0], res[2][b])


def test_bf16_store_matches_oracle_on_stored_values(pb, orc):
    """bf16 storage: the corpus IS the bf16-rounded values and the query is rounded
    to bf16 on entry (a bf16 index computes on bf16 operands on both sides); the
    oracle scans the same rounded values (fp32 accumulate, order v1/epc=8)."""
    n, dim, k = 20_000, 384, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    q = orc.synth_rows(2, 0, 0, 1, dim)[0]
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, k))
        back, _, _ = ix.get_rows(0, 100)
    stored = orc.round_bf16(rows)
    assert np.array_equal(back, stored[:100])
    want = orc.search(stored, ids, orc.round_bf16(q), k, mode=orc.MODE_F32_V1, epc=8)
    assert_same_result(got, want, what="bf16")


def test_cosine_metric_unnormalised(pb, orc):
    """lib.rs:67-77 cosine on un-normalised rows, norms computed in-kernel."""
    n, dim, k = 10_000, 768, 10
    rows = orc.synth_rows(1, orc.DIST_SCALED, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    q = orc.synth_rows(2, orc.DIST_SCALED, 0, 1, dim)[0]
    with pb.Index(dim, metric=pb.PCV_METRIC_COSINE) as ix:
        ix.set_rows(rows, ids)
        got = _one(ix.search(q, k))
    want = orc.search(rows, ids, q, k, metric=orc.METRIC_COSINE, mode=orc.MODE_F32_V1)
    assert_same_result(got, want, what="cosine", cosine=True)
    truth = orc.np_search(rows, ids, q, k, metric=orc.METRIC_COSINE)
    assert np.array_equal(got[0], truth[0])
    np.testing.assert_allclose(got[2], truth[2], rtol=1e-5)


def test_prenormalise_flag(pb, orc):
    n, dim = 2000, 384
    rows = orc.synth_rows(1, orc.DIST_SCALED, 0, n, dim)
    with pb.Index(dim, flags=pb.PCV_FLAG_PRENORMALISE) as ix:
        ix.set_rows(rows, np.arange(n))
        back, _, _ = ix.get_rows(0, n)
    assert np.array_equal(back, orc.normalise_rows(rows))


def test_replace_source(pb, orc):
    """rebuild_source (search.rs:58-79): swap one source's rows, keep the rest."""
    dim = 384
    rows = orc.synth_rows(8, 0, 0, 900, dim)
    ids = np.arange(1, 901, dtype=np.int64)
    src = np.repeat([1, 2, 3], 300).astype(np.int64)
    new_rows = orc.synth_rows(9, 0, 0, 450, dim)
    new_ids = np.arange(2000, 2450, dtype=np.int64)
    q = new_rows[17]
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids, src)
        ix.replace_source(2, new_rows, new_ids)
        all_rows = np.concatenate([rows[:300], new_rows, rows[600:]])
        all_ids = np.concatenate([ids[:300], new_ids, ids[600:]])
        all_src = np.concatenate([src[:300], np.full(450, 2), src[600:]])
        for flt in (None, [2], [1, 3]):
            got = _one(ix.search(q, 10, sources=flt))
            want = orc.search(all_rows, all_ids, q, 10, source_ids=all_src, sources=flt, mode=orc.MODE_F32_V1)
            assert_same_result(got, want, what=f"replace flt={flt}")
        ix.replace_source(7, new_rows[:5], new_ids[:5] + 10_000)  # new source appended
        got = _one(ix.search(q, 3, sources=[7]))
        assert int(got[3]) == 3
        ix.replace_source(1, np.zeros((0, dim), np.float32), np.zeros(0, np.int64))  # removal
        got = _one(ix.search(q, 3, sources=[1]))
        assert int(got[3]) == 0
        assert ix.stats().n_rows == 450 + 300 + 5


def test_concurrent_searches_one_handle(pb, orc):
    """Searcher: Send + Sync (app_state.rs:75): barrier-started threads share one
    handle, in the style of the reference's batch_sender.rs:187-221 test."""
    import threading
    n, dim, k = 50_000, 384, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = orc.synth_rows(2, 0, 0, 8, dim)
    want = [orc.search(rows, ids, qs[i], k, mode=orc.MODE_F32_V1) for i in range(8)]
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        barrier = threading.Barrier(8)
        errs = []

        def work(i):
            try:
                barrier.wait()
                for _ in range(5):
                    assert_same_result(_one(ix.search(qs[i], k)), want[i], what=f"thread {i}")
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert not errs, errs


def test_logical_shards_merge_invariance(pb, orc):
    """Shard-count invariance without NCCL: G row-range shards on one device, local
    top-k each, merged by the K5 kernel == the unsharded result, for G in 1,2,4,8."""
    import torch
    n, dim, k, B = 40_000, 384, 10, 3
    rows = orc.synth_rows(1, 0, 0, n, dim)
    rows[n // 2 + 5] = rows[3]  # a duplicate straddling shards
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = np.stack([rows[3], orc.synth_rows(2, 0, 0, 1, dim)[0], rows[n - 1]])
    with pb.Index(dim) as full:
        full.set_rows(rows, ids)
        ref = full.search(qs, k)
    for G in (1, 2, 4, 8):
        bounds = np.linspace(0, n, G + 1).astype(int)
        sims = torch.empty((G, B, k), dtype=torch.float32, device="cuda")
        cids = torch.empty((G, B, k), dtype=torch.int64, device="cuda")
        shards = []
        for g in range(G):
            ix = pb.Index(dim)
            ix.set_rows(rows[bounds[g]:bounds[g + 1]], ids[bounds[g]:bounds[g + 1]])
            r = ix.search(qs, k)
            s = np.where(r[0] >= 0, r[2], -np.inf).astype(np.float32)
            i = np.where(r[0] >= 0, r[0], np.iinfo(np.int64).max)
            sims[g] = torch.from_numpy(s).cuda()
            cids[g] = torch.from_numpy(i).cuda()
            shards.append(ix)
        o_ids = torch.empty((B, k), dtype=torch.int64, device="cuda")
        o_scores = torch.empty((B, k), dtype=torch.float32, device="cuda")
        o_sims = torch.empty((B, k), dtype=torch.float32, device="cuda")
        o_cnt = torch.empty(B, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        shards[0].merge_candidates_device(sims.data_ptr(), cids.data_ptr(), G, B, k, o_ids.data_ptr(),
                                          o_scores.data_ptr(), o_sims.data_ptr(), o_cnt.data_ptr())
        shards[0].synchronize()
        assert np.array_equal(o_ids.cpu().numpy(), ref[0]), f"G={G}"
        assert np.array_equal(o_sims.cpu().numpy(), ref[2]), f"G={G}"
        assert np.array_equal(o_scores.cpu().numpy(), ref[1]), f"G={G}"
        for ix in shards:
            ix.close()


def test_zero_norm_rows_and_queries_rejected_under_cosine(pb):
    """lib.rs:67-77 divides by the L2 norm with no epsilon (a zero row would be NaN); the
    library refuses zero-norm rows at load and zero-norm queries at search (PCV_ERR_ZERO_NORM)."""
    dim = 64
    rows = np.eye(8, dim, dtype=np.float32)
    rows[3] = 0.0
    with pb.Index(dim, metric=pb.PCV_METRIC_COSINE) as ix:
        with pytest.raises(pb.PcvError) as e:
            ix.set_rows(rows, np.arange(8))
        assert e.value.code == 8
        rows[3, 5] = 2.0
        ix.set_rows(rows, np.arange(8))
        with pytest.raises(pb.PcvError) as e:
            ix.search(np.zeros(dim, np.float32), 3)
        assert e.value.code == 8
        ids, scores, sims, cnt = ix.search(rows[3], 3)
        assert ids[0, 0] == 3 and abs(sims[0, 0] - 1.0) < 1e-6 and ids[0, 1] == 5  # e5 shares the direction
        # finite but tiny / huge: the fp32 sum of squares underflows to 0 or overflows — NaN or 0 similarities
        # would follow silently, so these are refused like a zero vector
        for bad in (1e-25, 1e25):
            with pytest.raises(pb.PcvError) as e:
                ix.search(np.full(dim, bad, np.float32), 3)
            assert e.value.code == 8
            r2 = rows.copy()
            r2[2] = bad
            with pytest.raises(pb.PcvError) as e:
                ix.set_rows(r2, np.arange(8))
            assert e.value.code == 8
    # the dot metric has no norm: zero rows are legal there
    with pb.Index(dim) as ix:
        rows[3] = 0.0
        ix.set_rows(rows, np.arange(8))
        assert int(ix.search(rows[0], 8)[3][0]) == 8


def test_single_row_sources_and_like_lookup(pb, orc):
    """Sources holding one row each (per-source top-k of the reference degenerates to that row),
    and the `--like ID` flow of perceive-cli/cmd/search.rs:64-85: fetch a stored row, search with it."""
    n, dim = 40, 384
    rows = orc.synth_rows(4, 0, 0, n, dim)
    ids = np.arange(100, 100 + n, dtype=np.int64)
    src = np.arange(n, dtype=np.int64)  # every row its own source
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids, src)
        got = ix.search(rows[7], 5, sources=[7, 9, 11])
        want = orc.search(rows, ids, rows[7], 5, source_ids=src, sources=[7, 9, 11], mode=orc.MODE_F32_V1)
        assert int(got[3][0]) == 3 and np.array_equal(got[0][0][:3], want[0])
        stored, sid, ssrc = ix.get_rows(7, 1)  # --like: the stored embedding of item 107
        assert sid[0] == 107 and ssrc[0] == 7 and np.array_equal(stored[0], rows[7])
        like = ix.search(stored[0], 3)
        assert like[0][0][0] == 107  # an item is its own nearest neighbour


@pytest.mark.parametrize("dim,nq", [(384, 7), (768, 5), (100, 3)])
def test_bf16_small_batch_scan_matches_oracle(pb, orc, dim, nq):
    """Batches below the tensor-path threshold run the scan with 2 queries per pass on bf16 rows:
    every query bit-identical to the oracle's v1-order scan of the same bf16 operands."""
    n, k = 12_000, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    with pb.Index(dim, store=pb.PCV_BF16) as ix:
        ix.set_rows(rows, ids)
        res = ix.search(qs, k)
        assert ix.stats().last_kernel == 1
    stored, qr = orc.round_bf16(rows), orc.round_bf16(qs)
    for b in range(nq):
        want = orc.search(stored, ids, qr[b], k, mode=orc.MODE_F32_V1, epc=8)
        assert_same_result(_one(res, b), want, what=f"bf16 batch query {b} dim={dim}")


def test_hidden_rows_are_cut_out_of_the_scan(pb, orc):
    """SURVEY.md 8 f1 (opt-in; the reference never reads `Searcher.hidden`, search.rs:34): with a
    hidden set installed the result is bit-identical to an index built without those rows."""
    dim, n, k = 384, 5000, 10
    rng = np.random.default_rng(9)
    rows = orc.synth_rows(8, 0, 0, n, dim)
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64)
    src = rng.integers(0, 3, size=n).astype(np.int64) * 5  # sources 0, 5, 10
    q = rows[123]
    order = np.lexsort((ids, src))  # storage order: (source, id)
    seg_last = [int(ids[order[np.nonzero(src[order] == s)[0][-1]]]) for s in (0, 5)]
    seg_first = [int(ids[order[np.nonzero(src[order] == s)[0][0]]]) for s in (5, 10)]
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids, src)
        base = _one(ix.search(q, k))
        assert base[0][0] == ids[123]
        hide = [int(i) for i in base[0][:3]]                       # the best hits
        hide += [int(ids[order[j]]) for j in (0, 1, 2, n - 1)]     # a run at the start, the very last row
        hide += seg_last + seg_first                               # both sides of segment boundaries
        hide += [10 ** 9, -4]                                      # ids that are not in the index
        hide += hide[:2]                                           # duplicates are harmless
        ix.set_hidden(hide)
        keep = ~np.isin(ids, hide)

        def check(what):
            for flt in (None, [5], [0, 10], []):
                got = _one(ix.search(q, k, sources=flt))
                want = orc.search(rows[keep], ids[keep], q, k, source_ids=src[keep], sources=flt, mode=orc.MODE_F32_V1)
                assert_same_result(got, want, what=f"{what} filter={flt}")
                assert not set(got[0][: int(got[3])].tolist()) & set(hide)
        check("hidden")
        # small batch through the scan kernel too
        qs = orc.synth_rows(12, 0, 0, 3, dim)
        res = ix.search(qs, k)
        for b in range(3):
            want = orc.search(rows[keep], ids[keep], qs[b], k, mode=orc.MODE_F32_V1)
            assert_same_result(_one(res, b), want, what=f"hidden batch q{b}")
        # the set outlives a rebuild of one source (rows move, ids stay hidden)
        m5 = src == 5
        ix.replace_source(5, rows[m5], ids[m5])
        check("after replace_source")
        # hiding a whole source, then everything
        ix.set_hidden(ids[src == 0])
        got = _one(ix.search(q, k))
        want = orc.search(rows[src != 0], ids[src != 0], q, k, mode=orc.MODE_F32_V1)
        assert_same_result(got, want, what="source 0 hidden")
        ix.set_hidden(ids)
        assert int(ix.search(q, k)[3][0]) == 0
        # n == 0 restores the reference behaviour
        ix.set_hidden([])
        c = int(base[3])
        assert_same_result(_one(ix.search(q, k)), (base[0][:c], base[1][:c], base[2][:c]), what="hidden cleared")


def test_find_id_and_embedding_of(pb, orc):
    """`--like ID` (perceive-cli/cmd/search.rs:64-85) served from the resident matrix."""
    dim, n = 384, 300
    rows = orc.synth_rows(4, 0, 0, n, dim)
    ids = np.arange(1000, 1000 + n, dtype=np.int64)[::-1].copy()
    src = (np.arange(n) % 4).astype(np.int64)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids, src)
        for j in (0, 17, n - 1):
            assert np.array_equal(ix.embedding_of(int(ids[j])), rows[j])
            assert ix.get_rows(ix.find_id(int(ids[j])), 1)[1][0] == ids[j]
        assert ix.find_id(999) is None and ix.embedding_of(5) is None
    with pb.Index(dim) as ix:  # dense synthetic ids: id = first_row + row + 1
        ix.generate_synthetic(50, seed=1, first_row=100)
        assert ix.find_id(101) == 0 and ix.find_id(150) == 49 and ix.find_id(151) is None and ix.find_id(100) is None


def test_search_host_raw_pointers_and_result_paths_agree(pb, orc, monkeypatch):
    """pcv_search on caller-owned host buffers (what a Rust/C caller passes).  Small result sets are
    stored by the kernel straight into mapped pinned memory, larger ones come back through one copy:
    both must equal the oracle, and each other."""
    dim, n = 384, 9000
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = orc.synth_rows(2, 0, 0, 3, dim)
    with pb.Index(dim) as ix:
        ix.set_rows(rows, ids)
        for k in (10, 400):  # 3 x 10 x 16 B < 4 KB (zero-copy), 3 x 400 x 16 B > 4 KB (copy)
            o_ids = np.full((3, k), 7, dtype=np.int64)
            o_sc = np.zeros((3, k), dtype=np.float32)
            o_si = np.zeros((3, k), dtype=np.float32)
            o_c = np.zeros(3, dtype=np.uint32)
            ix.search_host(qs.ctypes.data, 3, k, o_ids.ctypes.data, o_sc.ctypes.data, o_si.ctypes.data, o_c.ctypes.data)
            for b in range(3):
                want = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
                assert_same_result((o_ids[b], o_sc[b], o_si[b], o_c[b]), want, what=f"search_host k={k} q{b}")
            monkeypatch.setenv("PCV_NO_ZERO_COPY_RESULTS", "1")
            again = ix.search(qs, k)
            monkeypatch.delenv("PCV_NO_ZERO_COPY_RESULTS")
            assert np.array_equal(again[0], o_ids) and np.array_equal(again[1], o_sc) and np.array_equal(again[3], o_c)
        # optional outputs may be omitted
        o_ids = np.empty((1, 5), dtype=np.int64)
        o_sc = np.empty((1, 5), dtype=np.float32)
        ix.search_host(qs.ctypes.data, 1, 5, o_ids.ctypes.data, o_sc.ctypes.data, 0, 0)
        assert np.array_equal(o_ids[0], orc.search(rows, ids, qs[0], 5, mode=orc.MODE_F32_V1)[0])


@pytest.mark.parametrize("store_name", ["f32", "split", "bf16"])
def test_repeated_single_query_searches_follow_every_change(pb, orc, store_name):
    """The reference's call — one query at a time, over and over on one Searcher — with everything that can change
    between calls changing: k, the source filter, the hidden set, one source's rows.  (Cached row ranges, padded
    query buffers and workspaces must follow; every search equals the oracle on the rows selected at that moment.)"""
    n, dim = 20_000, 384
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    src = (ids % 3).astype(np.int64)
    qs = orc.synth_rows(2, 0, 0, 12, dim)
    store = {"f32": pb.PCV_F32, "split": pb.PCV_F32_SPLIT, "bf16": pb.PCV_BF16}[store_name]
    stored, qq = (orc.round_bf16(rows), orc.round_bf16(qs)) if store_name == "bf16" else (rows, qs)
    epc = 8 if store_name == "bf16" else 4

    def want(b, k, sel=None):
        m = np.ones(n, bool) if sel is None else sel
        return orc.search(stored[m], ids[m], qq[b], k, mode=orc.MODE_F32_V1, epc=epc)

    with pb.Index(dim, store=store) as ix:
        ix.set_rows(rows, ids, src)
        for b in range(6):
            got = ix.search(qs[b], 10)
            assert_same_result(_one(got, 0), want(b, 10), what=f"{store_name} repeat {b}")
            assert ix.stats().last_kernel == 1 and ix.stats().last_launches >= 1
        for b in range(6, 9):
            assert_same_result(_one(ix.search(qs[b], 25), 0), want(b, 25), what="k changed")
        for b in range(3):
            assert_same_result(_one(ix.search(qs[b], 10, sources=[0, 2]), 0), want(b, 10, np.isin(src, [0, 2])), what="filtered")
        top = int(ix.search(qs[0], 10)[0][0][0])
        ix.set_hidden([top])
        for _ in range(3):
            assert_same_result(_one(ix.search(qs[0], 10), 0), want(0, 10, ids != top), what="hidden between searches")
        ix.set_hidden([])
        ix.replace_source(1, rows[:50], np.arange(900_001, 900_051, dtype=np.int64))
        keep = src != 1
        r2 = np.concatenate([stored[keep], stored[:50]])
        i2 = np.concatenate([ids[keep], np.arange(900_001, 900_051, dtype=np.int64)])
        for b in range(4):
            w = orc.search(r2, i2, qq[b], 10, mode=orc.MODE_F32_V1, epc=epc)
            assert_same_result(_one(ix.search(qs[b], 10), 0), w, what="rows replaced between searches")


def test_no_timing_flag_changes_nothing_but_the_counter(pcv_lib, orc):
    """PCV_FLAG_NO_TIMING: the same results, last_search_ms reads 0, every other counter still answers
    (the split filter's fallback count is read after a stream synchronisation instead of after the event)."""
    import perceive_b200 as pb
    n, dim, nq, k = 30_000, 384, 24, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    res = {}
    for flags in (0, pb.PCV_FLAG_NO_TIMING):
        with pb.Index(dim, store=pb.PCV_F32_SPLIT, flags=flags) as ix:
            ix.set_rows(rows, ids)
            one = ix.search(qs[:1], k)
            st1 = ix.stats()
            many = ix.search(qs, k)
            st = ix.stats()
        assert st1.last_kernel == 1 and st.last_kernel == 2 and st.last_launches >= 3
        assert st.last_fallback_queries == 0
        assert (st.last_search_ms == 0.0) == bool(flags), st.last_search_ms
        res[flags] = (one, many)
    for a, b in zip(res[0][0] + res[0][1], res[pb.PCV_FLAG_NO_TIMING][0] + res[pb.PCV_FLAG_NO_TIMING][1]):
        assert np.array_equal(a, b)
    with pytest.raises(pb.PcvError):
        pb.Index(dim, flags=4)
