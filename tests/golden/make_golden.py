"""Regenerates tests/golden/synth_fingerprint.json from the oracle's generator.
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
