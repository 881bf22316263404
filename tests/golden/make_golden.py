"""Regenerates tests/golden/synth_fingerprint.json and config1_top10.json from the oracle.
(known_answers.json is hand-written; it is NOT generated.)"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import oracle as orc  # noqa: E402

cases = []
for seed, dist, first_row, dim in [(1, 0, 0, 384), (2, 0, 0, 384), (1, 1, 7, 768), (5, 0, 123456789, 100)]:
    r = orc.synth_rows(seed, dist, first_row, 2, dim)
    cases.append({"seed": seed, "dist": dist, "first_row": first_row, "dim": dim,
                  "row0_first4_bits": r[0, :4].view(np.uint32).tolist(),
                  "row1_last2_bits": r[1, -2:].view(np.uint32).tolist()})
(Path(__file__).parent / "synth_fingerprint.json").write_text(json.dumps({"cases": cases}, indent=1) + "\n")
print("wrote synth_fingerprint.json")

# BASELINE configs[0] (the reference's own CPU-runnable case): 1 query vs 10k x 384 fp32, top-10.
# The oracle's result in the scan kernel's summation order, as raw bits: pins the oracle against
# regressions (test_oracle.py) and is what the CUDA path must reproduce (test_gpu_search.py).
n, dim, k, corpus_seed, query_seed = 10_000, 384, 10, 1, 2
rows = orc.synth_rows(corpus_seed, 0, 0, n, dim)
q = orc.synth_rows(query_seed, 0, 0, 1, dim)[0]
ids = np.arange(1, n + 1, dtype=np.int64)
w_ids, w_scores, w_sims = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1, epc=4)
twin = orc.np_search(rows, ids, q, k)
assert np.array_equal(w_ids, twin[0])
(Path(__file__).parent / "config1_top10.json").write_text(json.dumps({
    "workload": "BASELINE configs[0]: 1 query vs 10k x 384 fp32 unit-sphere rows, top-10",
    "rows": n, "dim": dim, "k": k, "corpus_seed": corpus_seed, "query_seed": query_seed,
    "ids": w_ids.tolist(),
    "sim_bits": np.asarray(w_sims, dtype=np.float32).view(np.uint32).tolist(),
    "score_bits": np.asarray(w_scores, dtype=np.float32).view(np.uint32).tolist(),
    "sims_float64": [float(x) for x in twin[2]],
}, indent=1) + "\n")
print("wrote config1_top10.json")
