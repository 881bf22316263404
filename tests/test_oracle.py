"""The oracle against its pins (CPU only): hand-derived known answers, the numpy
float64 twin, and the committed fingerprints of the synthetic generator."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = json.loads((Path(__file__).parent / "golden" / "known_answers.json").read_text())


def test_distance_known_answers(orc):
    for c in GOLDEN["distance"]:
        assert orc.distance_from_dot(c["dot"], c["len"]) == c["want"], c
        assert float(orc.np_distance(c["dot"], c["len"])) == c["want"], c


def test_codec_known_answers(orc):
    for c in GOLDEN["codec"]:
        blob = bytes.fromhex(c["hex"])
        assert orc.encode_embedding(c["floats"]) == blob
        assert orc.decode_embedding(blob).tolist() == c["floats"]
    for n in GOLDEN["codec_bad_lengths"]:
        with pytest.raises(ValueError):
            orc.decode_embedding(b"\0" * n)


def test_codec_roundtrip_random(orc):
    v = np.random.default_rng(0).standard_normal(768).astype(np.float32)
    assert np.array_equal(orc.decode_embedding(orc.encode_embedding(v)), v)
    assert orc.encode_embedding(v) == v.astype("<f4").tobytes()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_search_known_answers(orc, mode):
    g = GOLDEN["search_dim4"]
    rows, ids, src, q = np.array(g["rows"], np.float32), np.array(g["ids"]), np.array(g["source_ids"]), np.array(g["query"], np.float32)
    assert (rows @ q).tolist() == g["dots"]
    for c in g["cases"]:
        got = orc.search(rows, ids, q, c["k"], source_ids=src, sources=c["sources"], mode=mode)
        assert got[0].tolist() == c["ids"], c
        assert got[1].tolist() == c["scores"], c
        twin = orc.np_search(rows, ids, q, c["k"], source_ids=src, sources=c["sources"])
        assert twin[0].tolist() == c["ids"] and twin[1].tolist() == c["scores"], c


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_cosine_known_answers(orc, mode):
    g = GOLDEN["cosine_dim4"]
    rows, ids, q = np.array(g["rows"], np.float32), np.array(g["ids"]), np.array(g["query"], np.float32)
    got = orc.search(rows, ids, q, g["k"], metric=orc.METRIC_COSINE, mode=mode)
    assert got[0].tolist() == g["ids_want"]
    np.testing.assert_allclose(got[2], g["sims_want"], rtol=1e-6, atol=1e-7)
    twin = orc.np_search(rows, ids, q, g["k"], metric=orc.METRIC_COSINE)
    assert twin[0].tolist() == g["ids_want"]


def test_normalise_known_answers(orc):
    g = GOLDEN["normalise"]
    got = orc.normalise_rows(np.array(g["rows"], np.float32))
    np.testing.assert_allclose(got, np.array(g["want"], np.float32), rtol=1e-7, atol=0)
    assert not np.isnan(got).any()


@pytest.mark.parametrize("n,dim,k", [(2000, 384, 10), (500, 768, 20), (300, 100, 7), (1, 384, 3), (64, 1024, 64)])
def test_c_oracle_vs_numpy_twin(orc, n, dim, k):
    """All three C summation modes agree with the independent numpy float64 scan:
    identical ids, similarities within 1e-5 relative (north_star tolerance)."""
    rows = orc.synth_rows(11, 0, 0, n, dim)
    ids = np.random.default_rng(1).permutation(np.arange(1, n + 1))
    src = np.random.default_rng(2).integers(0, 3, n)
    q = orc.synth_rows(12, 0, 0, 1, dim)[0]
    twin = orc.np_search(rows, ids, q, k, source_ids=src, sources=[0, 2])
    for mode in (0, 1, 2):
        got = orc.search(rows, ids, q, k, source_ids=src, sources=[0, 2], mode=mode)
        assert np.array_equal(got[0], twin[0]), mode
        np.testing.assert_allclose(got[2], twin[2], rtol=1e-5, atol=1e-7)
        assert np.all(np.diff(got[1]) >= 0)


def test_dot_orders_close_to_f64(orc):
    rng = np.random.default_rng(3)
    for dim in (384, 768, 100, 1024, 17):
        a, b = rng.standard_normal(dim).astype(np.float32), rng.standard_normal(dim).astype(np.float32)
        truth = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
        for mode, epc in ((1, 4), (2, 4), (2, 8)):
            assert abs(orc.dot(a, b, mode, epc) - truth) <= 1e-5 * max(1.0, np.abs(a * b).sum())


def test_fast_baseline_agrees_with_oracle(orc):
    n, dim, k = 50_000, 384, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    q = orc.synth_rows(2, 0, 0, 1, dim)[0]
    want = orc.search(rows, np.arange(1, n + 1), q, k, mode=orc.MODE_F32_V1)
    for threads in (1, 3, 8):
        got = orc.search_fast(rows, q, k, threads=threads)
        assert np.array_equal(got[0], want[0])
        np.testing.assert_allclose(got[2], want[2], rtol=1e-5, atol=1e-7)


def test_synthetic_generator_fingerprint(orc):
    """Pins the generator (a change would silently change every benchmark corpus)."""
    fp = json.loads((Path(__file__).parent / "golden" / "synth_fingerprint.json").read_text())
    for c in fp["cases"]:
        r = orc.synth_rows(c["seed"], c["dist"], c["first_row"], 2, c["dim"])
        assert r[0, :4].view(np.uint32).tolist() == c["row0_first4_bits"]
        assert r[1, -2:].view(np.uint32).tolist() == c["row1_last2_bits"]
    r = orc.synth_rows(1, 0, 0, 4096, 384)
    np.testing.assert_allclose(np.linalg.norm(r.astype(np.float64), axis=1), 1.0, atol=2e-7)
    assert abs(r.mean()) < 1e-3 and abs(r.std() * np.sqrt(384) - 1.0) < 1e-3
    assert np.array_equal(orc.synth_rows(1, 0, 100, 5, 384), orc.synth_rows(1, 0, 0, 105, 384)[100:])


def test_bf16_rounding(orc):
    v = np.array([1.0, 1.00390625, 1.005859375, 1.001953125, -3.14159274, 65504.0], np.float32)
    got = orc.round_bf16(v)
    # 1+2^-8 is a tie between 1.0 and 1+2^-7: round to even (1.0); 1+3*2^-9 rounds up
    assert got[:4].tolist() == [1.0, 1.0, 1.0078125, 1.0]
    assert got[4] == np.float32(-3.140625) and got[5] == np.float32(65536.0)


def test_bootstrap_threshold_is_a_lower_bound_of_the_kth_best_score():
    """The invariant the batched path's bootstrap pass rests on: the k-th largest of
    the per-tile maximum scores over ANY set of >= k tiles never exceeds the k-th best score of the
    whole corpus — k maxima of k different tiles are k distinct rows — so filtering with `score >=
    threshold` cannot drop a member of the true top-k.  Checked on seeded random and adversarial
    score vectors (ties, one tile holding all the best rows, constant scores)."""
    rng = np.random.default_rng(0)
    tile = 64
    cases = []
    for n_tiles, k in ((40, 10), (512, 100), (130, 128), (16, 16)):
        s = rng.standard_normal(n_tiles * tile).astype(np.float32)
        cases.append((s, n_tiles, k))
        t = s.copy()
        t[:tile] += 10.0  # every top row inside tile 0: the bootstrap threshold is then loose, never wrong
        cases.append((t, n_tiles, k))
        cases.append((np.round(s * 2) / 2, n_tiles, k))  # heavy ties
        cases.append((np.full_like(s, 0.25), n_tiles, k))
    for scores, n_tiles, k in cases:
        kth_best = np.sort(scores)[::-1][k - 1]
        for boot_tiles in {k, min(n_tiles, 4 * k), n_tiles}:
            if boot_tiles < k or boot_tiles > n_tiles:
                continue
            maxima = scores[: boot_tiles * tile].reshape(boot_tiles, tile).max(axis=1)
            thr = np.sort(maxima)[::-1][k - 1]
            assert thr <= kth_best
            assert np.count_nonzero(scores >= thr) >= k
            survivors = np.sort(scores[scores >= thr])[::-1][:k]
            assert np.array_equal(survivors, np.sort(scores)[::-1][:k])


def test_oracle_against_blas_sdot_like_the_reference(orc):
    """The reference's score is `1 - ndarray_dot(a, b) / len`, clamped at 0 (search.rs:271-277), and
    ndarray's dot is BLAS `sdot` (perceive-core/Cargo.toml: ndarray `blas` feature).  numpy's fp32
    `rows @ q` is also BLAS (OpenBLAS sgemv here) — another vendor's summation order for the same
    arithmetic.  The restatement must agree with it within the north_star tolerance (1e-5 relative
    on the similarity), rank the same rows wherever fp32 BLAS itself separates them, and produce
    the same clamped distance from the same dot."""
    n, dim, k = 20_000, 384, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    for qi in range(4):
        q = orc.synth_rows(2, 0, qi, 1, dim)[0]
        blas = (rows @ q).astype(np.float32)  # sgemv, fp32 accumulate
        order = np.lexsort((ids, -blas))[:k + 1]
        got_ids, got_scores, got_sims = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1)
        np.testing.assert_allclose(got_sims.astype(np.float32), blas[got_ids - 1], rtol=1e-5, atol=1e-7)
        gaps = -np.diff(blas[order])
        if np.all(gaps > 2e-6):  # BLAS separates the top-(k+1): the ranking must be identical
            assert np.array_equal(got_ids, ids[order[:k]])
        # same dot -> same reference distance, bit for bit (fp32 divide, subtract, clamp)
        ref_dist = np.maximum(np.float32(1.0) - blas[got_ids - 1] / np.float32(dim), np.float32(0.0)).astype(np.float32)
        same_dot = got_sims.astype(np.float32) == blas[got_ids - 1]
        assert np.array_equal(got_scores[same_dot], ref_dist[same_dot])
        np.testing.assert_allclose(got_scores, ref_dist, rtol=0, atol=2e-7)
    # un-normalised rows with dot > len: the clamp collapses the distance to 0, order still by dot
    big = rows[:100] * 40.0
    q = rows[7] * 40.0
    got_ids, got_scores, got_sims = orc.search(big, ids[:100], q, 5, mode=orc.MODE_F32_V1)
    assert got_ids[0] == 8 and got_scores[0] == 0.0 and got_sims[0] > dim


def test_config1_golden_fixture(orc):
    """BASELINE configs[0] (1 query vs 10k x 384 fp32, top-10): the oracle reproduces the committed
    result bit for bit in the scan order, and the other summation orders and the float64 twin agree on
    the ids and within 1e-5 relative on the similarities."""
    g = json.loads((Path(__file__).parent / "golden" / "config1_top10.json").read_text())
    n, dim, k = g["rows"], g["dim"], g["k"]
    rows = orc.synth_rows(g["corpus_seed"], 0, 0, n, dim)
    q = orc.synth_rows(g["query_seed"], 0, 0, 1, dim)[0]
    ids = np.arange(1, n + 1, dtype=np.int64)
    w_ids, w_scores, w_sims = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1, epc=4)
    assert w_ids.tolist() == g["ids"]
    assert np.asarray(w_sims, dtype=np.float32).view(np.uint32).tolist() == g["sim_bits"]
    assert np.asarray(w_scores, dtype=np.float32).view(np.uint32).tolist() == g["score_bits"]
    for mode in (0, 2):
        o_ids, _, o_sims = orc.search(rows, ids, q, k, mode=mode)
        assert o_ids.tolist() == g["ids"]
        np.testing.assert_allclose(o_sims, g["sims_float64"], rtol=1e-5, atol=1e-7)
    f_ids, f_scores, f_sims = orc.search_fast(rows, q, k)  # the timed CPU baseline
    assert f_ids.tolist() == g["ids"]
    np.testing.assert_allclose(f_sims, g["sims_float64"], rtol=1e-5, atol=1e-7)


def test_reference_migration_fixtures_are_the_reference_bytes():
    """tests/golden/reference_migrations/*.sql are byte-for-byte copies of the reference's migrations
    (crates/perceive-core/migrations/); the fixture databases are built by executing them verbatim."""
    import hashlib
    from pathlib import Path
    want = {"00001_init.sql": "acdc16232be2be9884149d33cb4ce912ad1049e7eb71062781b087d25770fa10",
            "00002_tags.sql": "75e522023103091807d1602323d0b0dd734366a6ad9ad76df6b1095c8478ea05",
            "00003_model_7.sql": "07c8c6c3b18091702509569d2bf35645b5fbe7d04872dcc4e71405e6dc2aa5a4"}
    d = Path(__file__).parent / "golden" / "reference_migrations"
    for name, sha in want.items():
        assert hashlib.sha256((d / name).read_bytes()).hexdigest() == sha, name
    ref = Path("/root/reference/crates/perceive-core/migrations")
    if ref.exists():  # in the build container the originals are at hand: compare directly
        for name in want:
            assert (ref / name).read_bytes() == (d / name).read_bytes(), name
