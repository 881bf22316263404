#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cX] [--series S]

A "step" is ONE pass of the hot path over one batch of queries.  Two series:

  headline (one process, a one-GPU box, or --series headline): `value` is BASELINE
    config 2 — one query scored exactly against 1M x 384 fp32 rows, top-10 (K1 scan) — and
    the same JSON line carries a `workloads` object with full sub-records (ms_per_step,
    roofline, e2e, parity, clocks) for config 1 (10k rows, L2-resident), config 3 (1024 x 10M x
    384 bf16, top-100, tcgen05) and config 4 on ONE GPU (256 x 100M x 384 fp32-exact rows,
    top-10), plus `cpu_baseline` (config 2) and `cpu_baseline_c1` (config 1 exactly as named).
  scaling (under torchrun, or any launch on a box that shows more than one GPU, or
    --series scaling): every N — N = 1 included — runs BASELINE config 4, the config
    BASELINE names for 2/4/8 GPUs, row-sharded over the ranks (strong scaling), so that
    v_N / (N * v_1) compares like with like; at N = 8 a `workloads.c5` sub-record adds BASELINE
    config 5 (4096 x 50M x 768 bf16, cosine, top-50), which BASELINE names for 8 GPUs only.  Each step = local filter on the tensor cores +
    exact fp32 rescoring -> exchange of the (sim, id) candidates (stores into peer memory over
    NVLink, or ncclAllGather with --exchange nccl) -> merge; every rank holds the result.

Every record carries `value` (device-resident inputs, CUDA-event timed), `e2e` (host buffers
through the C-ABI `pcv_search`, copies inside the timed region), `roofline` for the dominant
kernel, and an untimed `parity` block (planted queries whose nearest neighbour is known a
priori + a float64 recomputation of returned similarities; at N > 1 also that every rank
holds the same result).  `--impl reference` times the CPU restatement of the reference's exact
scoring (oracle/baseline.c; the reference itself is Rust + hnsw_rs and cannot be built here) on a
bounded sample of the same workload and prints WHAT WAS TIMED; the figure scaled to the full
corpus sits under `cpu_baseline.extrapolated`.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "queries/sec (exact top-k cosine kNN)"
WORKLOADS = {
    # name: rows, dim, store, batch, k, metric, dist
    "c1": dict(rows=10_000, dim=384, store="f32", batch=1, k=10, metric="dot_ref", dist="unit_sphere",
               text="1 query vs 10kx384 fp32 docs, top-10 (BASELINE configs[0]; L2-resident)"),
    "c2": dict(rows=1_000_000, dim=384, store="f32", batch=1, k=10, metric="dot_ref", dist="unit_sphere",
               text="1 query vs 1Mx384 fp32 docs, top-10 (BASELINE configs[1])"),
    "c3": dict(rows=10_000_000, dim=384, store="bf16", batch=1024, k=100, metric="dot_ref", dist="unit_sphere",
               text="batch 1024 queries vs 10Mx384 bf16 docs, top-100 (BASELINE configs[2])"),
    "c4": dict(rows=100_000_000, dim=384, store="split", batch=256, k=10, metric="dot_ref", dist="unit_sphere",
               text="batch 256 queries vs 100Mx384 fp32 docs (exact fp32 values held as two 16-bit planes), top-10, "
                    "row-sharded (BASELINE configs[3])"),
    "c5": dict(rows=50_000_000, dim=768, store="bf16", batch=4096, k=50, metric="cosine", dist="scaled",
               text="batch 4096 queries vs 50Mx768 bf16 docs (distilbert-shaped), top-50, un-normalised rows, cosine; "
                    "1/|row| is computed on the device from the stored values by one pass at the first search and "
                    "cached, then applied in the GEMM epilogue (BASELINE configs[4])"),
}
CORPUS_SEED, QUERY_SEED = 1, 2

# Committed ncu summaries (one line per profiled launch: kernel, duration, dram bytes read / written ...)
# that `roofline.traffic` is read from at run time; written by tools/ncu_summary.py from the .ncu-rep of
# an `ncu --set full` capture of THIS script.  Newest round first; (workload, rows, file, kernel regex, note).
NCU_SUMMARIES = [
    ("c2", 1_000_000, ["profiles/r2_ncu_full_scan_c2.csv", "profiles/r1_ncu_full_scan_c2_final.csv"], r"scan_kernel",
     "scan_kernel, one launch = one step"),
    ("c3", 10_000_000, ["profiles/r2_ncu_full_gemm_c3.csv", "profiles/r1_ncu_full_gemm_c3_pair.csv"], r"gemm_topk_pair_kernel",
     "main pass of gemm_topk_pair_kernel (the bootstrap and threshold passes read the first 8192 tiles)"),
    ("c4", 12_500_000, ["profiles/r2_ncu_full_filter_c4_shard.csv"], r"gemm_topk_pair_kernel",
     "main pass of the hi-plane filter over a 12.5M-row shard: what one rank of the 8-GPU run launches"),
]


def _bytes_of(cell: str) -> float:
    m = re.match(r"\s*([0-9.eE+-]+)\s*([KMGT]?)byte", cell)
    if not m:
        return float("nan")
    return float(m.group(1)) * {"": 1.0, "K": 1e3, "M": 1e6, "G": 1e9, "T": 1e12}[m.group(2)]


def ncu_traffic(workload: str, rows: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, read from the newest
    committed ncu summary for exactly this workload and this many rows on the GPU; None when there is none."""
    for wl, n, files, kernel_re, note in NCU_SUMMARIES:
        if wl != workload or n != rows:
            continue
        for f in files:
            p = ROOT / f
            if not p.exists():
                continue
            best = None
            with open(p, newline="") as fh:
                for row in csv.DictReader(fh):
                    if not re.search(kernel_re, row.get("kernel", "")):
                        continue
                    b = _bytes_of(row.get("dram__bytes_read.sum", "")) + _bytes_of(row.get("dram__bytes_write.sum", ""))
                    if b == b and (best is None or b > best):  # the longest (main) pass of a multi-pass search
                        best = b
            if best is not None:
                return {"bytes": best, "source": f"{f} ({note})"}
    return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d["bf16_tflops"]), tf_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 100 ms for the life of the process (started long before
    any timed region: nvidia-smi's own start-up — NVML initialisation — perturbs a GPU that is being timed);
    `window(t0, t1)` summarises the samples that arrived during a timed region (or, for a region shorter than the
    sampling period, the ones closest to it)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples: list[tuple[float, float, float, list[str]]] = []  # (arrival time, sm MHz, max MHz, active reasons)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, smax = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.samples.append((time.monotonic(), sm, smax, [nm for nm, v in zip(self.NAMES, f[5:9]) if v.lower().startswith("active")]))

    def wait_ready(self, timeout_s: float = 3.0):
        """Block until the first sample has arrived, i.e. nvidia-smi's start-up is over."""
        deadline = time.monotonic() + timeout_s
        while self.proc is not None and not self.samples and time.monotonic() < deadline:
            time.sleep(0.02)

    def window(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.monotonic() + 0.25
        while (not self.samples or self.samples[-1][0] < t1) and time.monotonic() < deadline:
            time.sleep(0.02)  # let the sample that covers the end of the region arrive
        inside = [s for s in self.samples if t0 <= s[0] <= t1 + 0.11]  # a sample reports the 100 ms before it
        if not inside and self.samples:
            inside = sorted(self.samples, key=lambda s: abs(s[0] - t1))[:2]
        sm = [s[1] for s in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((s[2] for s in inside), default=None),
                "reasons": sorted({r for s in inside for r in s[3]}), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()


def host_threads() -> int:
    """Every host core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers; the
    CPU arm runs on rank 0 alone while the other ranks are idle, so it takes all cores regardless."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def config_of(w, world: int, exchange: str | None = None) -> dict:
    """The `config` object — the same function for both arms, so they print identical dicts."""
    esz = 2 if w["store"] == "bf16" else 4
    corpus = w["rows"] * w["dim"] * esz
    return {"workload": w["text"], "rows": w["rows"], "dim": w["dim"], "k": w["k"], "batch": w["batch"],
            "n_gpus": world, "sharding": f"rows/{world}" if world > 1 else "none",
            "l2": (f"corpus ({corpus / 1e9:.2f} GB) larger than L2 (126 MB); a fresh query batch every step"
                   if corpus > (126 << 20) else "corpus is L2-resident (smaller than 126 MB): not an HBM number"),
            "corpus_seed": CORPUS_SEED, "query_seed": QUERY_SEED}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_sample(w, max_rows: int, n_queries: int):
    from oracle import oracle as orc
    n_sample = min(w["rows"], max_rows)
    d_id = 1 if w["dist"] == "scaled" else 0
    rows = orc.synth_rows(CORPUS_SEED, d_id, 0, n_sample, w["dim"])
    qs = orc.synth_rows(QUERY_SEED, d_id, 0, n_queries, w["dim"])
    if w["store"] == "bf16":
        rows, qs = orc.round_bf16(rows), orc.round_bf16(qs)
    if w["metric"] == "cosine":  # norms taken out of the timed loop: the scan then ranks exactly by cosine
        rows = rows / np.linalg.norm(rows, axis=1, keepdims=True)
    return orc, rows, qs, n_sample


def cpu_scan_baseline(w, budget_s: float = 10.0, max_queries: int = 200):
    """`cpu_baseline` of a record: the oracle's CPU scan (oracle/baseline.c, kind 'port') on this box's
    host cores on a BOUNDED sample: at most 1M rows of the same synthetic corpus.  Single-query
    workloads run one query at a time as the reference's search_vector does (search.rs:157); batched
    workloads run the blocked batch pass.  Reports what was timed; the figure scaled linearly in the
    row count to the full corpus is `extrapolated`."""
    B = w["batch"]
    threads = host_threads()
    if B == 1:
        orc, rows, qs, n_sample = _cpu_sample(w, 1_000_000, max_queries)
        for i in range(3):
            orc.search_fast(rows, qs[i], w["k"], threads=threads)
        t0 = time.perf_counter()
        n = 0
        while n < max_queries and (n < 10 or time.perf_counter() - t0 < budget_s):
            orc.search_fast(rows, qs[n], w["k"], threads=threads)
            n += 1
        dt = time.perf_counter() - t0
        qps = n / dt
        sample = f"{n} queries, one at a time, each a full scan of {n_sample}x{w['dim']} fp32 rows of the corpus"
    else:
        orc, rows, qs, n_sample = _cpu_sample(w, 1_000_000, B)
        orc.search_batch_fast(rows[: n_sample // 8], qs, w["k"], threads=threads)
        t0 = time.perf_counter()
        n = 0
        while n < 1 or (time.perf_counter() - t0 < budget_s and n < 20):
            orc.search_batch_fast(rows, qs, w["k"], threads=threads)
            n += 1
        dt = time.perf_counter() - t0
        qps = n * B / dt
        sample = (f"{n} passes of the whole batch ({B} queries, blocked sgemm-style) over {n_sample}x{w['dim']} fp32 rows "
                  "of the corpus")
    out = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample,
           "sample_rows": n_sample, "ms_per_query": 1e3 / qps}
    if n_sample != w["rows"]:
        scale = w["rows"] / n_sample
        out["extrapolated"] = {"value": qps / scale, "unit": "queries/s",
                               "how": f"the sample's time x{scale:g} (the scan is linear in the row count)"}
    return out


def run_reference(args, w):
    """--impl reference: the reference's exact scoring on the host CPU (oracle port).  Each step is the
    workload's query batch against a bounded row sample; `value` is WHAT WAS TIMED."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    B, k = w["batch"], w["k"]
    threads = host_threads()
    total = args.steps + args.warmup
    if B == 1:
        orc, rows, qs, n_sample = _cpu_sample(w, 1_000_000, total)

        def step(i):
            orc.search_fast(rows, qs[i], k, threads=threads)
    else:
        # sized so that steps + warmup end within a few minutes on 8-16 cores (~0.25 TFLOP/s fp32)
        max_rows = int(max(50_000, min(1_000_000, 60 * 0.25e12 / (2.0 * B * w["dim"] * max(total, 1)))))
        orc, rows, qs, n_sample = _cpu_sample(w, max_rows, B * min(total, 4))
        qs = qs.reshape(-1, B, w["dim"])

        def step(i):
            orc.search_batch_fast(rows, qs[i % qs.shape[0]], k, threads=threads)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    step_s = dt / args.steps
    qps = B / step_s
    sample = (f"{args.steps} steps, each the workload's batch of {B} quer{'y' if B == 1 else 'ies'} "
              f"({'one scan per query' if B == 1 else 'one blocked sgemm-style pass'}) over {n_sample}x{w['dim']} fp32 rows"
              + ("" if n_sample == w["rows"] else f" — a SAMPLE of the {w['rows']}-row corpus; `value` is what was timed"))
    cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample, "sample_rows": n_sample}
    if n_sample != w["rows"]:
        scale = w["rows"] / n_sample
        cpu["extrapolated"] = {"value": qps / scale, "unit": "queries/s",
                               "how": f"the sample's time x{scale:g} (the scan is linear in the row count)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(w, world),
        "cpu_baseline": cpu,
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bf16_round(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return r.view(np.float32)


class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the search path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            self.sampler.start()
            self.sampler.wait_ready()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def parity_block(ctx: Ctx, ix, w, dist_id: int, host_rows) -> dict:
    """Untimed check of the index the timed region just used.  Planted queries: query j is corpus row
    p(j) (regenerated on the host by the library's generator, which is bit-identical to the device's),
    so its nearest neighbour is known a priori whatever the corpus size; then a float64 recomputation of
    the returned similarities from the returned ids, sortedness, counts, and — across ranks — that every
    rank holds the same result."""
    rows, dim, k, B = w["rows"], w["dim"], w["k"], w["batch"]
    nq = max(B, 8)
    pos = (np.arange(nq, dtype=np.int64) * 7919 + 13) % rows
    qs = np.concatenate([host_rows(int(p), 1) for p in pos])
    ids, scores, sims, cnt = ix.search(qs, k)
    ok_top1 = bool(np.array_equal(ids[:, 0], pos + 1))
    ok_cnt = bool(np.all(cnt == k))
    ok_sorted = bool(np.all(np.diff(sims, axis=1) <= 0))
    max_err, n_spot, ok_spot = 0.0, 0, True
    for b in np.unique(np.linspace(0, nq - 1, 6).astype(int)):
        got_rows = np.concatenate([host_rows(int(i) - 1, 1) for i in ids[b]])
        q = qs[b]
        if w["store"] == "bf16":
            got_rows, q = bf16_round(got_rows), bf16_round(q)
        t = got_rows.astype(np.float64) @ q.astype(np.float64)
        if w["metric"] == "cosine":
            t = t / np.linalg.norm(got_rows.astype(np.float64), axis=1) / np.linalg.norm(q.astype(np.float64))
        err = np.abs(sims[b].astype(np.float64) - t)
        max_err = max(max_err, float(err.max()))
        ok_spot = ok_spot and bool(np.all(err <= 1e-5 * np.abs(t) + 2e-6))
        n_spot += k
    ranks_agree = True
    if ctx.world > 1:
        torch = ctx.torch
        digest = torch.tensor([int(np.bitwise_xor.reduce((ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).ravel()) >> np.uint64(1))],
                              dtype=torch.int64, device=ctx.dev)
        allv = [torch.empty_like(digest) for _ in range(ctx.world)]
        ctx.dist.all_gather(allv, digest)
        ranks_agree = all(int(v.item()) == int(digest.item()) for v in allv)
    ok = ok_top1 and ok_cnt and ok_sorted and ok_spot and ranks_agree
    return {"checked": int(nq), "ok": bool(ok), "planted_top1": ok_top1, "counts": ok_cnt, "sorted": ok_sorted,
            "similarities_recomputed_in_float64": n_spot, "max_abs_err": max_err, "tolerance": "1e-5*|sim| + 2e-6",
            "ranks_agree": ranks_agree,
            "what": "planted queries (a corpus row is its own nearest neighbour) + float64 dot of the returned rows, untimed"}


def run_workload(ctx: Ctx, name: str, w: dict, args, steps: int, warmup: int) -> dict | None:
    """One workload on this process's GPU (its shard of the rows when world > 1).  Returns the record on
    rank 0, None elsewhere."""
    torch, dist = ctx.torch, ctx.dist
    import perceive_b200 as pb
    from perceive_b200 import _ffi
    from perceive_b200.distributed import attach_shard, shard_rows
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    rows, dim, k, B = w["rows"], w["dim"], w["k"], w["batch"]
    store = {"f32": pb.PCV_F32, "bf16": pb.PCV_BF16, "split": pb.PCV_F32_SPLIT}[w["store"]]
    esz = 2 if w["store"] == "bf16" else 4
    r0, r1 = shard_rows(rows, rank, world)
    metric = pb.PCV_METRIC_COSINE if w["metric"] == "cosine" else pb.PCV_METRIC_DOT_REF
    dist_id = pb.PCV_DIST_SCALED if w["dist"] == "scaled" else pb.PCV_DIST_UNIT_SPHERE
    # the library's own per-search CUDA timing events (pcv_stats.last_search_ms) are left out unless asked for: this
    # file times with its own events, and two event records per search are stream bubbles of their own
    # (tools/timing_events_probe.sh: config 1 19.5 -> 14.4 us per search on the device, config 2 227.3 -> 221.9 us)
    ix = pb.Index(dim, device=ctx.local_rank, store=store, metric=metric, flags=0 if args.timing_events else pb.PCV_FLAG_NO_TIMING)
    ix.generate_synthetic(r1 - r0, CORPUS_SEED, dist=dist_id, first_row=r0)
    exchange_used = attach_shard(ix, dist, rank, world, device=dev, exchange=args.exchange, max_records=max(B * k, 8 * k, 1 << 12))
    lib = _ffi.load()

    def host_rows(first: int, n: int, seed: int = CORPUS_SEED) -> np.ndarray:
        out = np.empty((n, dim), dtype=np.float32)
        _ffi.check(lib.pcv_synthetic_rows_host(seed, dist_id, first, n, dim, out.ctypes.data))
        return out

    total = steps + warmup
    # queries: the same synthetic stream on every rank.  A fresh batch every step, drawn round-robin
    # from a pool of distinct batches (pool bounded at ~64 MB; the library keeps no state between searches).
    pool = max(1, min(total, (64 << 20) // (B * dim * 4)))
    q_host = host_rows(0, pool * B, QUERY_SEED).reshape(pool, B, dim)

    # ---------------- device-resident arm (`value`) ----------------------------
    stream = torch.cuda.Stream(device=dev)  # the library launches on it; the events below are recorded on it
    assert stream.cuda_stream != 0
    ix.set_stream(stream.cuda_stream)
    d_q = torch.from_numpy(q_host).to(dev)
    o_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    o_scores = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_sims = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_cnt = torch.empty(B, dtype=torch.int32, device=dev)

    def step_device(i):
        ix.search_device(d_q[i % pool].data_ptr(), B, k, o_ids.data_ptr(), o_scores.data_ptr(), o_sims.data_ptr(),
                         o_cnt.data_ptr())

    h_ids = np.empty((B, k), dtype=np.int64)
    h_scores = np.empty((B, k), dtype=np.float32)
    h_sims = np.empty((B, k), dtype=np.float32)
    h_counts = np.empty(B, dtype=np.uint32)
    q_ptrs = [q_host[i].ctypes.data for i in range(pool)]
    o_ptrs = (h_ids.ctypes.data, h_scores.ctypes.data, h_sims.ctypes.data, h_counts.ctypes.data)

    torch.cuda.synchronize()
    # A GPU that has just been handed a fresh corpus needs ~30 ms of work before it runs at its steady rate (measured
    # on config 2, tools/cold_start_probe.py: consecutive 20-step regions right after the build take 0.250, 0.248,
    # 0.244, 0.243, 0.240, 0.238 ms per step).  Throughput is a steady-state figure, so the W warm-up steps are
    # preceded by untimed steps for at least 60 ms; both are reported.
    # (Every search of a sharded index is collective, so every rank must run the SAME number of them: the count is
    # derived on rank 0 from two timed steps and broadcast.)
    t_pre = time.perf_counter()
    for i in range(2):  # the first searches allocate workspaces and upload row ranges
        step_device(i)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for i in range(2):
        step_device(2 + i)
    torch.cuda.synchronize()
    per_step = max((time.perf_counter() - t2) / 2, 1e-6)
    n_more = int(min(400, max(0, np.ceil(0.06 / per_step))))
    if world > 1:
        t = torch.tensor([n_more], dtype=torch.int64, device=dev)
        dist.broadcast(t, 0)
        n_more = int(t.item())
    for i in range(n_more):
        step_device(4 + i)
    torch.cuda.synchronize()
    pre_steps = 4 + n_more
    pre_ms = 1e3 * (time.perf_counter() - t_pre)
    for i in range(warmup):
        step_device(i)
    ctx.barrier()
    t_region0 = time.monotonic()
    # The two arms are interleaved in blocks (device-resident block, then host-buffer block, ...) so that
    # both see the same clocks: a dense tensor step runs under a moving power cap.
    # (single-query steps are a quarter of a millisecond at full clocks: one block, or every block's idle-to-busy
    # transition — tens of microseconds — would be a tenth of what is timed)
    n_blocks = 1 if (steps < 8 or B == 1) else 4
    dev_ms, e2e_s, done = 0.0, 0.0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for blk in range(n_blocks):
        nb = steps * (blk + 1) // n_blocks - steps * blk // n_blocks
        ix.set_stream(stream.cuda_stream)
        ctx.barrier()
        ev0.record(stream)
        for i in range(nb):
            step_device(warmup + done + i)
        ev1.record(stream)
        ctx.barrier()
        dev_ms += ctx.max_over_ranks(ev0.elapsed_time(ev1))
        if blk == n_blocks - 1:
            st = ix.stats()
            last_ids = o_ids.cpu().numpy().copy()
        # end-to-end arm: pcv_search on caller-owned HOST buffers, exactly what a Rust/C caller passes:
        # every step copies that step's queries host->device and its results device->host
        ix.set_stream(None)
        if blk == 0:
            for i in range(min(warmup, 3)):
                ix.search_host(q_ptrs[i % pool], B, k, *o_ptrs)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(nb):
            ix.search_host(q_ptrs[(warmup + done + i) % pool], B, k, *o_ptrs)
        dt = time.perf_counter() - t0
        ctx.barrier()
        e2e_s += ctx.max_over_ranks(dt)
        done += nb
    clocks = ctx.sampler.window(t_region0, time.monotonic()) if rank == 0 else None
    assert np.array_equal(h_ids, last_ids), "device-resident and host-buffer arms disagree"
    launches_per_step = st.last_launches

    parity = parity_block(ctx, ix, w, dist_id, host_rows)
    rec = None
    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = dev_ms / steps
        qps = steps * B / (dev_ms * 1e-3)
        local_rows = r1 - r0
        local_bytes = local_rows * dim * esz  # algorithmic bytes one pass streams (N*d*sizeof)
        k2 = st.last_kernel == 2
        if k2 and w["store"] == "split":
            # config 4.  SURVEY 8d: "HBM-bound if an fp32-exact tensor path sustains >= 1 PF-equivalent, otherwise
            # compute-bound; report both".  ncu shows the tensor pipe 90-97.5 % active in the main pass, so the tensor
            # roofline is the bound reported as `frac`; the HBM view (SURVEY's algorithmic N*d*4 bytes, of which the
            # filter streams only the 2-byte hi plane) is given beside it.
            flops = 2.0 * B * local_rows * dim
            achieved = flops / (ms_per_step * 1e-3) / 1e12
            streamed = local_rows * dim * 2
            hbm_alg = local_bytes / (ms_per_step * 1e-3) / 1e9
            roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tf_sus"], "traffic": None, "peak_source": peaks["src"] + " (sustained cuBLAS bf16)",
                    "kernel": "pcv::gemm_topk_pair_kernel<6,SHAPE_BF16> over the hi plane (tcgen05.mma.cta_group::2 M256 N256 K16, bf16 -> f32 TMEM) "
                              "+ pcv::rescore_exact_kernel (fp32, K1's summation order)",
                    "flops_per_step": flops, "frac_of_burst_peak": achieved / peaks["tf"], "frac_of_nominal_2250TF": achieved / 2250.0,
                    "hbm_view": {"algorithmic_bytes_per_step": local_bytes, "algorithmic_GBps": hbm_alg,
                                 "algorithmic_frac_of_hbm_peak": hbm_alg / peaks["hbm"],
                                 "streamed_bytes_per_step": streamed, "streamed_GBps": streamed / (ms_per_step * 1e-3) / 1e9,
                                 "streamed_frac_of_hbm_peak": streamed / (ms_per_step * 1e-3) / 1e9 / peaks["hbm"],
                                 "note": "SURVEY 8d's algorithmic bytes are N*d*4; the filter reads N*d*2 of them (the hi plane) and "
                                         "the exact rescoring gathers 48 rows per query, so the algorithmic fraction can exceed 1"}}
        elif k2:  # tensor-bound: 2*B*N*d flops per step on this rank's shard
            flops = 2.0 * B * local_rows * dim
            achieved = flops / (ms_per_step * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tf_sus"], "traffic": None, "peak_source": peaks["src"] + " (sustained cuBLAS bf16)",
                    "kernel": ("pcv::gemm_topk_pair_kernel<6,SHAPE_BF16> (tcgen05.mma.cta_group::2 M256 N256 K16, bf16 -> f32 TMEM)" if dim <= 384 else
                               "pcv::gemm_topk_pair_kernel<12,SHAPE_WIDE> (tcgen05.mma.cta_group::2 M256 N128 K16, bf16 -> f32 TMEM, cosine epilogue)"),
                    "flops_per_step": flops, "frac_of_burst_peak": achieved / peaks["tf"],
                    "frac_of_nominal_2250TF": achieved / 2250.0,
                    "hbm_GBps_algorithmic": local_bytes / (ms_per_step * 1e-3) / 1e9}
        else:
            passes = (B + 3) // 4 if B > 1 else 1  # K1 scores up to 4 queries per pass over the rows
            achieved = passes * local_bytes / (ms_per_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm"], "traffic": None, "peak_source": peaks["src"],
                    "kernel": ("pcv::scan_kernel<SplitF32,...> (both 16-bit planes)" if w["store"] == "split" else
                               "pcv::scan_kernel<float,12,1,1,false>" if (esz == 4 and B == 1 and dim == 384) else
                               f"pcv::scan_kernel<{'float' if esz == 4 else 'bf16'},...>"), "bytes_per_launch": local_bytes,
                    "launches_per_step": passes, "frac_of_nominal_8TBs": achieved / 8000.0,
                    "peak_note": "the measured peak is a COPY (read + write bytes of b.copy_(a)); a read-only stream pays no "
                                 "write turn-arounds and can pass it: frac > 1 is not an error"}
        tr = ncu_traffic(name, local_rows)
        if tr:
            roof["traffic"] = tr["bytes"]
            roof["traffic_source"] = tr["source"]
        cfg = config_of(w, world)
        rec = {
            "metric": METRIC, "value": qps, "unit": "queries/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": w["store"], "data": "synthetic",
            "config": cfg,
            "e2e": {"value": steps * B / e2e_s, "unit": "queries/s",
                    "h2d_bytes_per_step": B * dim * 4, "d2h_bytes_per_step": B * k * 16 + B * 4,
                    "ms_per_step": 1e3 * e2e_s / steps},
            "gpu_launches": int(launches_per_step) * steps,
            "launches_per_step": int(launches_per_step),
            "library_timing_events": bool(args.timing_events),  # PCV_FLAG_NO_TIMING unless --timing-events
            "timing": f"{n_blocks} interleaved blocks of device-resident steps (CUDA events) and host-buffer steps (wall clock), max over ranks",
            "pre_warmup": {"steps": pre_steps, "ms": round(pre_ms, 1),
                           "why": "untimed steps for >= 60 ms before the W warm-up steps: a GPU just handed a fresh corpus takes ~30 ms to reach its steady rate"},
            "exchange": (("stores into peer memory over NVLink + epoch flags, merged in the same launch"
                          if exchange_used == "p2p" else "ncclAllGather + merge kernel") if world > 1 else "none"),
            "roofline": roof,
            "parity": parity,
            "clocks": clocks,
        }
    ix.close()
    del d_q, o_ids, o_scores, o_sims, o_cnt
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    from perceive_b200 import _build
    _build.build()  # no-op when the in-tree .so is current; raises if it cannot be built
    ctx = Ctx()
    torch = ctx.torch
    series = args.series
    if series == "auto":
        under_torchrun = "TORCHELASTIC_RUN_ID" in os.environ or "RANK" in os.environ
        series = "scaling" if (ctx.world > 1 or under_torchrun or torch.cuda.device_count() > 1) else "headline"
    if args.workload is not None:
        names, extra = [args.workload], []
    elif series == "scaling":
        # config 4 at every N; at 8 GPUs — the only place BASELINE names it — config 5 rides along as a sub-record
        names, extra = ["c4"], (["c5"] if (ctx.world == 8 and not args.no_workloads) else [])
    else:
        names, extra = ["c2"], ([] if args.no_workloads else ["c1", "c3", "c4"])

    def steps_for(w):
        return args.steps if args.steps is not None else (200 if w["batch"] == 1 else 20)

    def load(name):
        w = dict(WORKLOADS[name])
        if args.rows is not None and name == names[0]:
            w["text"] += f" [rows overridden: {args.rows} instead of {w['rows']}]"
            w["rows"] = args.rows
        if args.store is not None and name == names[0]:
            w["text"] += f" [storage overridden: {args.store} instead of {w['store']}]"
            w["store"] = args.store
        return w

    w = load(names[0])
    out = run_workload(ctx, names[0], w, args, steps_for(w), args.warmup)
    if ctx.rank == 0:
        out["series"] = (f"{series}: " + ("every N runs BASELINE config 4 (strong scaling over row shards)" if series == "scaling" and args.workload is None
                                         else "BASELINE config 2 is `value`; configs 3 and 4 (one GPU) are under `workloads`" if args.workload is None
                                         else f"--workload {args.workload}"))
    subs = {}
    for nm in extra:
        sw = load(nm)
        try:
            rec = run_workload(ctx, nm, sw, args, steps_for(sw), args.warmup)
        except Exception as e:  # noqa: BLE001 - the headline line must still be printed
            rec = {"error": f"{type(e).__name__}: {e}"}
        subs[nm] = rec
    if ctx.rank == 0:
        if subs:
            out["workloads"] = subs
        if ctx.world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_scan_baseline(w)
            if args.workload is None and series == "headline":
                c1 = cpu_scan_baseline(dict(WORKLOADS["c1"]), budget_s=3.0, max_queries=2000)
                c1["workload"] = WORKLOADS["c1"]["text"]
                out["cpu_baseline_c1"] = c1
        print(json.dumps(out), flush=True)
    # the JSON line is out: never let teardown (CUDA context, NCCL, helper threads) stall the run
    wd = threading.Timer(30.0, lambda: os._exit(0))
    wd.daemon = True
    wd.start()
    ctx.sampler.stop()
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="run exactly this workload (default: chosen by --series)")
    ap.add_argument("--series", default="auto", choices=["auto", "headline", "scaling"],
                    help="headline: c2 as `value` + c3/c4 sub-records (one GPU); scaling: c4 at every N.  auto = scaling "
                         "under torchrun or on a box showing more than one GPU, else headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timing-events", action="store_true", help="keep the library's per-search CUDA timing events (pcv_stats.last_search_ms)")
    ap.add_argument("--no-workloads", action="store_true", help="headline series without the c3 / c4 sub-records")
    ap.add_argument("--rows", type=int, default=None, help="override the workload's corpus size (experiments only)")
    ap.add_argument("--store", default=None, choices=["f32", "bf16", "split"],
                    help="override the workload's storage type (experiments only), e.g. config 2 on the two-plane fp32 layout")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how shards exchange their top-k candidates")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        # the same workload choice as the GPU arm, without touching CUDA
        name = args.workload
        if name is None:
            series = args.series
            if series == "auto":
                multi = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) > 1 or "TORCHELASTIC_RUN_ID" in os.environ or "RANK" in os.environ
                if not multi:  # the same question the GPU arm asks torch: how many GPUs does this process see?
                    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                    if vis is not None:
                        multi = len([d for d in vis.split(",") if d.strip()]) > 1
                    else:
                        try:
                            out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
                            multi = sum(1 for ln in out.splitlines() if ln.startswith("GPU ")) > 1
                        except Exception:
                            multi = False
                series = "scaling" if multi else "headline"
            name = "c4" if series == "scaling" else "c2"
        w = dict(WORKLOADS[name])
        if args.rows is not None:
            w["text"] += f" [rows overridden: {args.rows} instead of {w['rows']}]"
            w["rows"] = args.rows
        if args.steps is None:
            args.steps = 200 if w["batch"] == 1 else 20
        run_reference(args, w)
    else:
        run_ours(args)


if __name__ == "__main__":
    if os.environ.get("BENCH_WATCHDOG"):  # debugging aid: dump every thread's stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    main()
    sys.stdout.flush()
    sys.stderr.flush()
