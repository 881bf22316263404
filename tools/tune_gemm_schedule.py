"""Sweep the K2/K3 pass schedule on a B200 (development aid, not product).

    python tools/tune_gemm_schedule.py [--rows 10000000] [--dim 384] [--store bf16] [--batch 1024] [--k 100]

The library reads PCV_GEMM_FIRST_TILES / PCV_GEMM_PASS_RATIO / PCV_GEMM_DENSE_TILES / PCV_GEMM_BOOT_TILES / PCV_GEMM_AFTER_BOOT_TILES / PCV_NO_PDL from the
environment at every search, so one resident index serves every configuration; configurations are
visited round-robin (`--rounds`) so clock drift under the power cap hits them alike.  One JSON line
per configuration: median / min ms per batch over all rounds (CUDA events on the search stream)."""
import argparse
import json
import os
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import perceive_b200 as pb  # noqa: E402

CONFIGS = [  # (first tiles, ratio, dense tiles, bootstrap tiles [0 = no bootstrap pass], first pass after the bootstrap ends at
             #  [0 = bootstrap x ratio; huge = ONE pass over everything], PCV_NO_PDL)
    (32, 4, 8192, 512, 0, 0),   # the default
    (32, 4, 8192, 512, 0, 1),   # ... without programmatic dependent launch
    (32, 4, 8192, 1024, 0, 0),
    (32, 4, 8192, 2048, 0, 0),
    (32, 4, 8192, 2048, 8192, 0),        # bootstrap, [0, 8192), rest
    (32, 4, 8192, 2048, 1 << 30, 0),     # bootstrap, then one pass
    (32, 4, 8192, 4096, 1 << 30, 0),
    (32, 4, 8192, 1024, 1 << 30, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--store", default="bf16", choices=["bf16", "split"])
    ap.add_argument("--metric", default="dot", choices=["dot", "cosine"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ix = pb.Index(a.dim, store=pb.PCV_BF16 if a.store == "bf16" else pb.PCV_F32_SPLIT,
                  metric=pb.PCV_METRIC_COSINE if a.metric == "cosine" else pb.PCV_METRIC_DOT_REF)
    ix.generate_synthetic(a.rows, 1, dist=pb.PCV_DIST_SCALED if a.metric == "cosine" else pb.PCV_DIST_UNIT_SPHERE)
    stream = torch.cuda.Stream(device=dev)
    ix.set_stream(stream.cuda_stream)
    B, k = a.batch, a.k
    n_q = a.iters + 2
    q = torch.randn((n_q, B, a.dim), device=dev, dtype=torch.float32)
    q = q / q.norm(dim=-1, keepdim=True)
    o_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    o_sc = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_si = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_c = torch.empty(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def run(i):
        ix.search_device(q[i % n_q].data_ptr(), B, k, o_ids.data_ptr(), o_sc.data_ptr(), o_si.data_ptr(), o_c.data_ptr())

    times = {c: [] for c in CONFIGS}
    launches = {}
    ref_ids = None
    for rnd in range(a.rounds):
        for cfg in CONFIGS:
            (os.environ["PCV_GEMM_FIRST_TILES"], os.environ["PCV_GEMM_PASS_RATIO"], os.environ["PCV_GEMM_DENSE_TILES"],
             os.environ["PCV_GEMM_BOOT_TILES"], os.environ["PCV_GEMM_AFTER_BOOT_TILES"], os.environ["PCV_NO_PDL"]) = map(str, cfg)
            run(0)  # warm-up, and the same batch for every configuration: results must not depend on the schedule
            stream.synchronize()
            launches[cfg] = int(ix.stats().last_launches)
            if ref_ids is None:
                ref_ids = o_ids.clone()
            elif not torch.equal(ref_ids, o_ids):
                print(json.dumps({"config": cfg, "error": "result differs from the default schedule"}), flush=True)
            for i in range(a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                run(1 + i)
                e1.record(stream)
                e1.synchronize()
                times[cfg].append(e0.elapsed_time(e1))
    for cfg in CONFIGS:
        t = times[cfg]
        print(json.dumps({"first": cfg[0], "ratio": cfg[1], "dense": cfg[2], "boot": cfg[3], "after_boot": cfg[4], "no_pdl": cfg[5], "ms_median": round(statistics.median(t), 4),
                          "ms_min": round(min(t), 4), "launches": launches[cfg], "n": len(t)}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
