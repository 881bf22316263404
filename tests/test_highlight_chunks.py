"""Highlighter chunk scoring (crates/perceive-core/model/highlight.rs:103-127):
oracle known answers on the CPU, the CUDA entry point against the oracle on the GPU."""
import numpy as np
import pytest


def test_oracle_best_chunks_known_answers(orc):
    q = np.array([1.0, 0.0, 2.0], dtype=np.float32)
    chunks = np.array([[1, 0, 0],    # doc 0: scores 1, 5, 5   -> last maximum = 2
                       [1, 9, 2],
                       [3, -4, 1],
                       [0, 0, -1],   # doc 2 (doc 1 is empty): scores -2, -3 -> 0
                       [1, 1, -2],
                       [2, 0, 2]],   # doc 3: single chunk -> 0
                      dtype=np.float32)
    best, scores = orc.np_best_chunks(q, chunks, [3, 3, 5, 6])
    assert scores.tolist() == [1, 5, 5, -2, -3, 6]
    assert best.tolist() == [2, -1, 0, 0]
    best, scores = orc.np_best_chunks(q, np.zeros((0, 3), np.float32), [0, 0])
    assert best.tolist() == [-1, -1] and scores.size == 0


def test_best_chunks_argument_validation_without_device(pcv_lib):
    assert pcv_lib.pcv_index_best_chunks(None, None, None, 0, None, 1, None, None, None) == 1


@pytest.mark.gpu
def test_best_chunks_matches_oracle(pcv_lib, orc):
    import perceive_b200 as pb
    rng = np.random.default_rng(21)
    for dim in (384, 768, 100):
        q = orc.synth_rows(2, 0, 0, 1, dim)[0]
        counts = [5, 0, 1, 37, 64, 0, 3, 130]          # chunks per document; empty documents included
        ends = np.cumsum(counts).astype(np.uint32)
        chunks = orc.synth_rows(5, 0, 0, int(ends[-1]), dim)
        chunks[2] = chunks[4]                            # equal scores inside doc 0, the later one the maximum?
        chunks[4] = chunks[2] = q                        # ... make both THE maximum: last one (4) must win
        chunks[6 + 10] = chunks[6 + 30]                  # a tie that is not the maximum changes nothing
        with pb.Index(dim) as ix:
            best, best_score, scores = ix.best_chunks(q, chunks, ends)
            assert ix.stats().last_launches == 1
            w_best, w_scores = orc.np_best_chunks(q, chunks, ends)
            np.testing.assert_allclose(scores, w_scores, rtol=1e-5, atol=1e-6)
            assert best[0] == 4 and best[1] == -1 and best[5] == -1 and best[2] == 0
            start = 0
            for d, end in enumerate(ends):
                sl = w_scores[start:end]
                if sl.size:
                    # the chosen chunk is a maximum up to fp32 rounding, and it is exactly the GPU's own maximum
                    assert sl[best[d]] >= sl.max() - (1e-5 * abs(sl.max()) + 1e-6)
                    g = scores[start:end]
                    assert best[d] == g.size - 1 - int(np.argmax(g[::-1])) and best_score[d] == g.max()
                    if np.sort(sl)[-1] - (np.sort(sl)[-2] if sl.size > 1 else -np.inf) > 1e-4:
                        assert best[d] == w_best[d]
                else:
                    assert best[d] == -1 and best_score[d] == 0.0
                start = int(end)
            # no documents / no chunks at all
            b0, _, s0 = ix.best_chunks(q, np.zeros((0, dim), np.float32), [0, 0, 0])
            assert b0.tolist() == [-1, -1, -1] and s0.size == 0
            # errors: NaN (the reference panics, highlight.rs:124), boundaries that are not cumulative
            bad = chunks.copy()
            bad[3, 1] = np.nan
            with pytest.raises(pb.PcvError) as e:
                ix.best_chunks(q, bad, ends)
            assert e.value.code == 3
            with pytest.raises(pb.PcvError) as e:
                ix.best_chunks(q, chunks, [5, 3])
            assert e.value.code == 1
            with pytest.raises(pb.PcvError):
                ix.best_chunks(q, chunks[:4], [5])
