"""Shared helpers for the parity tests (GPU path vs oracle)."""
import numpy as np


def assert_same_result(got, want, *, bitexact_sims=True, rtol=1e-5, atol=1e-7, what="", cosine=False):
    """got = (ids[k], scores[k], sims[k], count) from the C ABI (one query);
    want = (ids[c], scores[c], sims[c]) from the oracle."""
    g_ids, g_scores, g_sims, g_cnt = got
    w_ids, w_scores, w_sims = want
    assert int(g_cnt) == len(w_ids), f"{what}: count {g_cnt} != {len(w_ids)}"
    c = int(g_cnt)
    assert np.array_equal(g_ids[:c], w_ids), f"{what}: ids differ\n got  {g_ids[:c]}\n want {w_ids}"
    if bitexact_sims:
        assert np.array_equal(g_sims[:c].astype(np.float32), np.asarray(w_sims, dtype=np.float32)), \
            f"{what}: sims not bit-identical\n got  {g_sims[:c]}\n want {w_sims}"
        assert np.array_equal(g_scores[:c], w_scores), f"{what}: scores not bit-identical"
    else:
        np.testing.assert_allclose(g_sims[:c], w_sims, rtol=rtol, atol=atol, err_msg=what)
    # unused tail slots
    assert np.all(g_ids[c:] == -1), f"{what}: unused id slots must be -1"
    assert np.all(np.isinf(g_scores[c:])), f"{what}: unused score slots must be +inf"
    if cosine:  # score == similarity: best first means descending
        assert np.all(np.diff(g_scores[:c]) <= 0), f"{what}: cosine scores not descending"
    else:  # reference distance: ascending (the reference sorts ascending by score, search.rs:179)
        assert np.all(np.diff(g_scores[:c]) >= 0), f"{what}: scores not ascending"
