"""recall@k of the bf16 and fp32-split indexes against the fp32 exact path (K1 on fp32 rows, which is
bit-identical to the oracle), on the same synthetic corpus — the figure north_star asks to see beside
the bf16 tolerance.  A property of the STORAGE choice, not a parity claim: a bf16 index answers exactly
for the bf16-rounded values (tests/test_gpu_gemm.py); this measures how often rounding the values
moves a document across the k-th place.

    python tools/recall_vs_fp32.py [--rows 1000000] [--queries 256]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

import perceive_b200 as pb  # noqa: E402
from perceive_b200 import _ffi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--queries", type=int, default=256)
    a = ap.parse_args()
    q = np.empty((a.queries, a.dim), dtype=np.float32)
    _ffi.check(_ffi.load().pcv_synthetic_rows_host(2, 0, 0, a.queries, a.dim, q.ctypes.data))
    out = {"rows": a.rows, "dim": a.dim, "queries": a.queries}
    for k in (10, 100):
        with pb.Index(a.dim, store=pb.PCV_F32) as ix:
            ix.generate_synthetic(a.rows, seed=1)
            exact = ix.search(q, k)[0]
        for name, store in (("bf16", pb.PCV_BF16), ("split", pb.PCV_F32_SPLIT)):
            with pb.Index(a.dim, store=store) as ix:
                ix.generate_synthetic(a.rows, seed=1)
                got = ix.search(q, k)[0]
                assert ix.stats().last_kernel == 2
            hit = sum(len(set(e.tolist()) & set(g.tolist())) for e, g in zip(exact, got))
            same_order = float(np.mean([np.array_equal(e, g) for e, g in zip(exact, got)]))
            out[f"recall@{k}_{name}"] = hit / (a.queries * k)
            out[f"identical_lists@{k}_{name}"] = same_order
    print(json.dumps(out))


if __name__ == "__main__":
    main()
