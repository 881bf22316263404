"""Sweep the K1 scan's launch knobs on a B200 (development aid, not product).

    python tools/tune_scan.py [--rows 1000000] [--dim 384] [--store f32] [--iters 50]

Knobs are read by the library from the environment at every launch
(PCV_SCAN_TILE_BYTES, PCV_SCAN_NSLOTS, PCV_SCAN_L2HINT, PCV_SCAN_NB)."""
import argparse
import itertools
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import perceive_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--store", default="f32")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--tiles", default="3072,6144,12288")
    ap.add_argument("--slots", default="2,3,4,6,8")
    ap.add_argument("--hints", default="0,1")
    ap.add_argument("--lprs", default="3")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ix = pb.Index(a.dim, store=pb.PCV_F32 if a.store == "f32" else pb.PCV_BF16)
    ix.generate_synthetic(a.rows, 1)
    stream = torch.cuda.Stream(device=dev)
    ix.set_stream(stream.cuda_stream)
    B, k = a.batch, a.k
    q = torch.randn((a.iters + 5, B, a.dim), device=dev, dtype=torch.float32)
    q = q / q.norm(dim=-1, keepdim=True)
    o_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    o_sc = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_si = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_c = torch.empty(B, dtype=torch.int32, device=dev)
    esz = 4 if a.store == "f32" else 2
    nbytes = a.rows * ((a.dim * esz + 15) // 16 * 16)
    torch.cuda.synchronize()
    results = []
    for lpr, tile, slots, hint in itertools.product(a.lprs.split(","), a.tiles.split(","), a.slots.split(","), a.hints.split(",")):
        os.environ["PCV_SCAN_LPR_LOG2"] = lpr
        os.environ["PCV_SCAN_TILE_BYTES"] = tile
        os.environ["PCV_SCAN_NSLOTS"] = slots
        os.environ["PCV_SCAN_L2HINT"] = hint
        try:
            for i in range(5):
                ix.search_device(q[i].data_ptr(), B, k, o_ids.data_ptr(), o_sc.data_ptr(), o_si.data_ptr(), o_c.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(a.iters):
                ix.search_device(q[5 + i].data_ptr(), B, k, o_ids.data_ptr(), o_sc.data_ptr(), o_si.data_ptr(), o_c.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            st = ix.stats()
            r = dict(lpr=int(lpr), tile=int(tile), slots=int(slots), hint=int(hint), ms=round(ms, 4),
                     GBps=round(st.last_scan_bytes / ms / 1e6, 1), launches=st.last_launches)
        except pb.PcvError as e:
            r = dict(lpr=int(lpr), tile=int(tile), slots=int(slots), hint=int(hint), error=e.message)
        results.append(r)
        print(json.dumps(r), flush=True)
    ok = [r for r in results if "ms" in r]
    if ok:
        best = min(ok, key=lambda r: r["ms"])
        print("BEST", json.dumps(best), "bytes", nbytes)


if __name__ == "__main__":
    main()
