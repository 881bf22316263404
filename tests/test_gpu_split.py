"""Parity of K3 — fp32-accurate batched search on the tensor cores (PCV_F32_SPLIT rows:
hi/lo bf16 planes, hi*hi + hi*lo + lo*hi per K step) — against the float64 dot of the
ORIGINAL fp32 rows and queries.  north_star tolerance: 1e-5 relative in fp32; the
dropped lo*lo term and the 2^-17 storage residual are an absolute error floor on
near-zero sums, hence the small atol.  Ranking: exact top-k outside the tie band."""
import numpy as np
import pytest

from test_gpu_gemm import check_batch

pytestmark = pytest.mark.gpu

SPLIT_RTOL = 1e-5
SPLIT_ATOL = 2e-6


@pytest.fixture(scope="module")
def pb(pcv_lib):
    import perceive_b200
    return perceive_b200


@pytest.mark.parametrize("n,dim,nq,k", [(40_000, 384, 1, 10), (40_000, 384, 5, 10), (25_000, 384, 200, 10),
                                         (30_000, 128, 64, 100), (10_000, 100, 33, 7), (300, 384, 3, 50)])
def test_split_rows_match_fp32_truth(pb, orc, n, dim, nq, k):
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(rows, ids)
        back, _, _ = ix.get_rows(0, min(n, 256))
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    # stored value = hi + lo: within 2^-17 relative of the fp32 original
    assert np.all(np.abs(back - rows[:back.shape[0]]) <= np.abs(rows[:back.shape[0]]) * 2.0 ** -16 + 1e-30)
    err = check_batch(res, rows, ids, qs, k, what=f"split n={n} dim={dim} B={nq} k={k}", rtol=SPLIT_RTOL, atol=SPLIT_ATOL)
    print(f"K3 max |sim - f64(fp32 inputs)| = {err:.3e}")


def test_split_unnormalised_rows_and_sources(pb, orc):
    """Un-normalised rows (dot > dim clamps the reference distance to 0, order still by
    similarity) and a source filter, on split rows."""
    n, dim, nq, k = 20_000, 384, 40, 10
    rows = orc.synth_rows(1, orc.DIST_SCALED, 0, n, dim) * 8.0
    qs = orc.synth_rows(2, orc.DIST_SCALED, 0, nq, dim) * 8.0
    ids = np.arange(1, n + 1, dtype=np.int64)
    src = (np.arange(n) // 7000).astype(np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(rows, ids, src)
        for flt in (None, [0, 2], [1]):
            res = ix.search(qs, k, sources=flt)
            mask = None if flt is None else np.isin(src, flt)
            check_batch(res, rows, ids, qs, k, selected=mask, what=f"split sources={flt}", rtol=SPLIT_RTOL, atol=2e-4)
        assert res[1].min() == 0.0  # clamp reached (search.rs:277)


def test_split_rejects_what_it_does_not_implement(pb):
    with pytest.raises(pb.PcvError) as e:
        pb.Index(384, store=pb.PCV_F32_SPLIT, metric=pb.PCV_METRIC_COSINE)
    assert e.value.code == 5
    with pytest.raises(pb.PcvError):
        pb.Index(768, store=pb.PCV_F32_SPLIT)
    with pb.Index(384, store=pb.PCV_F32_SPLIT) as ix:
        ix.set_rows(np.eye(384, dtype=np.float32), np.arange(384))
        with pytest.raises(pb.PcvError) as e:
            ix.search(np.ones(384, np.float32), 200)
        assert e.value.code == 5


def test_split_config4_shape_subsample(pb, orc):
    """BASELINE config 4's shape (batch 256, 384-d fp32-accurate, top-10) on 300k
    device-generated rows."""
    n, dim, nq, k = 300_000, 384, 256, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    qs = orc.synth_rows(2, 0, 0, nq, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    with pb.Index(dim, store=pb.PCV_F32_SPLIT) as ix:
        ix.generate_synthetic(n, seed=1)
        res = ix.search(qs, k)
        st = ix.stats()
    assert st.last_kernel == 2
    err = check_batch(res, rows, ids, qs, k, what="config4-shape", rtol=SPLIT_RTOL, atol=SPLIT_ATOL)
    # recall@k against the exact fp32-order oracle (ids identical unless an epsilon-tie)
    same = 0
    for b in range(0, nq, 16):
        w_ids, _, _ = orc.search(rows, ids, qs[b], k, mode=orc.MODE_F32_V1)
        same += int(np.array_equal(res[0][b], w_ids))
    print(f"K3 config-4 shape: max err {err:.3e}; {same}/16 sampled queries have ids identical to the fp32 scan; "
          f"{st.last_launches} launches, {st.last_search_ms:.3f} ms")
