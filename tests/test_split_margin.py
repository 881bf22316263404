"""The mathematics behind K3's completeness proof (perceive_b200/csrc/pcv_rescore.cuh), checked on the CPU in
float64: for fp32 rows x held as hi = top 16 bits (x truncated to bf16) + lo, and queries rounded to bf16 (RNE),

    | q.x - bf16(q).hi(x) |  <=  |q - bf16(q)| * |x|  +  |bf16(q)| * |x - hi(x)|          (Cauchy-Schwarz)

which is what `split_query_margin_kernel` evaluates from the query and two per-index maxima.  A row outside the
filter's candidate set scores t <= t_kf on the tensor cores, hence at most t_kf + margin exactly; the test also
replays the whole filter -> rescore -> proof decision in numpy on a small corpus, adversarial inputs included,
and checks that whenever the proof holds the candidate set really contains the exact top-k."""
import numpy as np


def hi_plane(x: np.ndarray) -> np.ndarray:
    """fp32 -> the value its top 16 bits encode (truncation to bf16), as the load kernel stores it."""
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def lo_plane(x: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFF)).astype(np.uint16)


def bf16_rne(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).view(np.float32)


def margin(q: np.ndarray, rows: np.ndarray) -> float:
    """The device formula (data part + accumulation slack), in float64."""
    r64 = rows.astype(np.float64)
    xmax = np.sqrt((r64 ** 2).sum(axis=1).max())
    emax = np.sqrt(((r64 - hi_plane(rows).astype(np.float64)) ** 2).sum(axis=1).max())
    q64, qh = q.astype(np.float64), bf16_rne(q).astype(np.float64)
    data = np.linalg.norm(q64 - qh) * xmax + np.linalg.norm(qh) * emax
    return float(data + rows.shape[1] * 2.0 ** -22 * np.linalg.norm(q64) * xmax)


def test_planes_rebuild_the_fp32_value_exactly():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(10_000), [0.0, -0.0, 3.4e38, -1e-38, 1.17549435e-38, 65504.0]]).astype(np.float32)
    hi16 = (x.view(np.uint32) >> np.uint32(16)).astype(np.uint32)
    back = ((hi16 << np.uint32(16)) | lo_plane(x).astype(np.uint32)).view(np.float32)
    assert np.array_equal(back.view(np.uint32), x.view(np.uint32))
    # truncation never increases the magnitude and loses less than one bf16 ulp (2^-7 relative)
    h = hi_plane(x).astype(np.float64)
    assert np.all(np.abs(h) <= np.abs(x.astype(np.float64)))
    big = np.abs(x) > 1e-30
    assert np.all(np.abs(x[big].astype(np.float64) - h[big]) < 2.0 ** -7 * np.abs(x[big].astype(np.float64)))


def test_filter_error_is_within_the_margin_random_and_adversarial():
    rng = np.random.default_rng(1)
    dim = 384
    rows = rng.standard_normal((4_000, dim)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    # adversarial rows: every element sits just below the next bf16 value (maximal truncation error) ...
    worst = (hi_plane(rows[:500]).view(np.uint32) | np.uint32(0xFFFF)).view(np.float32)
    rows = np.concatenate([rows, worst, rows[:200] * 37.5])  # ... and un-normalised rows
    for trial in range(20):
        q = rng.standard_normal(dim).astype(np.float32)
        if trial % 4 == 1:  # a query aligned in sign with the truncation error of every worst-case row
            q = np.abs(q) * np.sign(worst[trial])
        if trial % 4 == 2:
            q *= 100.0
        m = margin(q, rows)
        exact = rows.astype(np.float64) @ q.astype(np.float64)
        filt = hi_plane(rows).astype(np.float64) @ bf16_rne(q).astype(np.float64)
        assert np.all(np.abs(exact - filt) <= m), (trial, np.abs(exact - filt).max(), m)


def _decide(rows, q, k, kf):
    """filter -> rescore -> proof, as the device does it (float64 stands in for both arithmetic paths)."""
    t = hi_plane(rows).astype(np.float64) @ bf16_rne(q).astype(np.float64)
    s = rows.astype(np.float64) @ q.astype(np.float64)
    cand = np.argsort(-t, kind="stable")[:kf]
    best = cand[np.argsort(-s[cand], kind="stable")[:k]]
    full = len(cand) == kf and kf < len(rows)
    proven = (not full) or (t[cand[-1]] + margin(q, rows) < s[best[-1]])
    return best, proven, np.argsort(-s, kind="stable")[:k]


def test_whenever_the_proof_holds_the_candidates_contain_the_exact_top_k():
    rng = np.random.default_rng(2)
    dim, k, kf = 384, 10, 48
    rows = rng.standard_normal((30_000, dim)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    proven_n = 0
    for _ in range(12):
        q = rng.standard_normal(dim).astype(np.float32)
        best, proven, truth = _decide(rows, q, k, kf)
        proven_n += proven
        if proven:
            assert np.array_equal(best, truth)
    assert proven_n >= 10, "on unit-sphere data the proof should hold for nearly every query"
    # all rows identical: every score ties, the candidate set can never be proven complete -> exact fallback
    same = np.tile(rows[:1], (2_000, 1))
    _, proven, _ = _decide(same, rng.standard_normal(dim).astype(np.float32), k, kf)
    assert not proven
    # a dense cluster of near-duplicates around the best row: more than kf rows within the margin -> fallback
    q = rows[0].copy()
    cluster = (rows[0][None, :] + 1e-4 * rng.standard_normal((200, dim))).astype(np.float32)
    _, proven, _ = _decide(np.concatenate([rows, cluster]), q, k, kf)
    assert not proven
