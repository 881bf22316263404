#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2]

A "step" is ONE pass of the hot path over one batch.  N=1 default: BASELINE
config 2, one query scored exactly against 1M x 384 fp32 rows, top-10 (K1 scan).
N>1 (torchrun, one rank per GPU) default: BASELINE config 4 — the config BASELINE
names for 2/4/8 GPUs — 256 queries against 100M x 384 fp32-accurate rows, top-10,
the corpus row-sharded over the ranks (strong scaling): each step = local tcgen05
search (K3) -> exchange of the (sim,id) candidates (stores into peer memory over
NVLink, or ncclAllGather with --exchange nccl) -> merge, every rank holding the
result.  `--workload c2 --gpus N` runs config 2 row-sharded instead (latency-bound).

The JSON line carries `value` (device-resident inputs, CUDA-event timed), `e2e`
(host buffers through the C-ABI `pcv_search`, copies inside the timed region),
`roofline` for the dominant kernel and `cpu_baseline` (the oracle's CPU scan timed
on this box's host cores).  `--impl reference` times that CPU scan as its own arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (rows, dim, store, batch, k, metric, dist)
    "c2": dict(rows=1_000_000, dim=384, store="f32", batch=1, k=10, metric="dot_ref", dist="unit_sphere",
               text="1 query vs 1Mx384 fp32 docs, top-10 (BASELINE configs[1])"),
    "c1": dict(rows=10_000, dim=384, store="f32", batch=1, k=10, metric="dot_ref", dist="unit_sphere",
               text="1 query vs 10kx384 fp32 docs, top-10 (BASELINE configs[0]; L2-resident)"),
    "c4": dict(rows=100_000_000, dim=384, store="split", batch=256, k=10, metric="dot_ref", dist="unit_sphere",
               text="batch 256 queries vs 100Mx384 fp32-accurate docs (hi/lo bf16 split rows), top-10, row-sharded (BASELINE configs[3])"),
    "c5": dict(rows=50_000_000, dim=768, store="bf16", batch=4096, k=50, metric="cosine", dist="scaled",
               text="batch 4096 queries vs 50Mx768 bf16 docs (distilbert-shaped), top-50, un-normalised rows, cosine with "
                    "row norms computed on the device, row-sharded (BASELINE configs[4])"),
    "c3": dict(rows=10_000_000, dim=384, store="bf16", batch=1024, k=100, metric="dot_ref", dist="unit_sphere",
               text="batch 1024 queries vs 10Mx384 bf16 docs, top-100 (BASELINE configs[2])"),
}
CORPUS_SEED, QUERY_SEED = 1, 2

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
# `ncu --set full` captures (profiles/r1_ncu_full_scan_c2_final.csv, profiles/r1_ncu_full_gemm_c3_pair.csv);
# only for the exact workloads those captures were taken on
NCU_TRAFFIC = {
    ("c2", 1_000_000): dict(bytes=1.536087e9 + 3.842e6, algorithmic=1.536e9,
                            source="profiles/r1_ncu_full_scan_c2_final.csv (scan_kernel, one launch = one step)"),
    ("c3", 10_000_000): dict(bytes=6.881363e9 + 17.49e6, algorithmic=69933 * 128 * 768.0,
                             source="profiles/r1_ncu_full_gemm_c3_pair.csv (main pass of gemm_topk_pair_kernel: "
                                    "69933 of the 78125 tiles; the bootstrap pass and the two short threshold passes read the other 8192, 512 of them twice)"),
}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d["bf16_tflops"]), tf_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines: list[str] = []
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
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """Every host core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers; the
    CPU arm runs on rank 0 alone while the other ranks are idle, so it takes all cores regardless."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_scan_baseline(w, budget_s: float = 12.0, max_queries: int = 200):
    """Time the oracle's CPU scan (oracle/baseline.c, kind 'port') on this box's host
    cores on a BOUNDED sample of the bench workload: at most 1M rows of the same
    synthetic corpus (bf16 workloads: the bf16-rounded values), one query at a time as
    the reference's search_vector does (search.rs:157).  Rows beyond the sample are
    accounted for by scaling the per-row time (the scan is linear in N).
    Returns (queries_per_s on the FULL corpus, cores, sample text, ms_per_query on the full corpus)."""
    from oracle import oracle as orc
    n_sample = min(w["rows"], 1_000_000)
    d_id = 1 if w["dist"] == "scaled" else 0
    rows = orc.synth_rows(CORPUS_SEED, d_id, 0, n_sample, w["dim"])
    qs = orc.synth_rows(QUERY_SEED, d_id, 0, max_queries, w["dim"])
    if w["store"] == "bf16":
        rows, qs = orc.round_bf16(rows), orc.round_bf16(qs)
    if w["metric"] == "cosine":  # norms taken out of the timed loop: the scan then ranks exactly by cosine
        rows = rows / np.linalg.norm(rows, axis=1, keepdims=True)
    threads = host_threads()
    for i in range(3):
        orc.search_fast(rows, qs[i], w["k"], threads=threads)
    t0 = time.perf_counter()
    n = 0
    while n < max_queries and (n < 10 or time.perf_counter() - t0 < budget_s):
        orc.search_fast(rows, qs[n], w["k"], threads=threads)
        n += 1
    dt = time.perf_counter() - t0
    scale = w["rows"] / n_sample
    sample = f"{n} queries, each a full scan of {n_sample}x{w['dim']} fp32 rows of the corpus"
    if scale != 1.0:
        sample += f"; time scaled x{scale:g} to the {w['rows']}-row corpus"
    return n / (dt * scale), threads, sample, 1e3 * dt * scale / n


def run_reference(args, w):
    """--impl reference: the reference's exact scoring on the host CPU (oracle
    port; the reference itself — Rust + hnsw_rs — cannot be built here).  Each step
    is one query batch of the workload on a bounded row sample (see cpu_scan_baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    n_sample = min(w["rows"], 1_000_000)
    bq = min(w["batch"], 8)  # queries actually timed per step
    d_id = 1 if w["dist"] == "scaled" else 0
    rows = orc.synth_rows(CORPUS_SEED, d_id, 0, n_sample, w["dim"])
    qs = orc.synth_rows(QUERY_SEED, d_id, 0, (args.steps + args.warmup) * bq, w["dim"])
    if w["store"] == "bf16":
        rows, qs = orc.round_bf16(rows), orc.round_bf16(qs)
    if w["metric"] == "cosine":  # norms taken out of the timed loop: the scan then ranks exactly by cosine
        rows = rows / np.linalg.norm(rows, axis=1, keepdims=True)
    threads = host_threads()
    for i in range(args.warmup * bq):
        orc.search_fast(rows, qs[i], w["k"], threads=threads)
    t0 = time.perf_counter()
    for i in range(args.steps * bq):
        orc.search_fast(rows, qs[args.warmup * bq + i], w["k"], threads=threads)
    dt = time.perf_counter() - t0
    scale = (w["rows"] / n_sample) * (w["batch"] / bq)  # to one full step of the workload
    step_s = dt / args.steps * scale
    qps = w["batch"] / step_s
    sample = (f"{args.steps} steps x {bq} queries, each a full scan of {n_sample}x{w['dim']} fp32 rows"
              + (f"; time scaled x{scale:g} to batch {w['batch']} x {w['rows']} rows" if scale != 1.0 else ""))
    print(json.dumps({
        "impl": "reference", "metric": "queries/sec (exact top-k cosine kNN)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": w["store"], "data": "synthetic",
        "config": {"workload": w["text"], "rows": w["rows"], "dim": w["dim"], "k": w["k"], "batch": w["batch"]},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    import perceive_b200 as pb
    from perceive_b200 import _build
    _build.build()  # no-op when the in-tree .so is current; raises if it cannot be built

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the search path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows, dim, k, B = w["rows"], w["dim"], w["k"], w["batch"]
    store = {"f32": pb.PCV_F32, "bf16": pb.PCV_BF16, "split": pb.PCV_F32_SPLIT}[w["store"]]
    esz = 2 if w["store"] == "bf16" else 4
    from perceive_b200.distributed import attach_shard, shard_rows
    r0, r1 = shard_rows(rows, rank, world)
    metric = pb.PCV_METRIC_COSINE if w["metric"] == "cosine" else pb.PCV_METRIC_DOT_REF
    dist_id = pb.PCV_DIST_SCALED if w["dist"] == "scaled" else pb.PCV_DIST_UNIT_SPHERE
    ix = pb.Index(dim, device=local_rank, store=store, metric=metric)
    ix.generate_synthetic(r1 - r0, CORPUS_SEED, dist=dist_id, first_row=r0)
    exchange_used = attach_shard(ix, dist, rank, world, device=dev, exchange=args.exchange, max_records=max(B * k, 1 << 12))

    total = args.steps + args.warmup
    # queries: the same synthetic stream on every rank (host generator of the library).
    # A fresh batch every step, drawn round-robin from a pool of distinct batches
    # (pool bounded at ~64 MB; the library keeps no state between searches).
    from perceive_b200 import _ffi
    pool = max(1, min(total, (64 << 20) // (B * dim * 4)))
    q_host = np.empty((pool * B, dim), dtype=np.float32)
    _ffi.check(_ffi.load().pcv_synthetic_rows_host(QUERY_SEED, dist_id, 0, pool * B, dim, q_host.ctypes.data))
    q_host = q_host.reshape(pool, B, dim)

    # ---------------- device-resident arm (`value`) ----------------------------
    # a dedicated (non-default) torch stream: the library launches on it and the
    # torch CUDA events below are recorded on it
    stream = torch.cuda.Stream(device=dev)
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

    torch.cuda.synchronize()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    st = ix.stats()
    launches_per_step = st.last_launches
    last_ids = o_ids.cpu().numpy().copy()
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())

    # ---------------- end-to-end arm (`e2e`): host buffers through pcv_search ----
    # pcv_search on caller-owned host buffers, exactly what a Rust/C caller passes (no per-call numpy
    # allocations): every step copies that step's queries host->device and its results device->host
    ix.set_stream(None)
    h_ids = np.empty((B, k), dtype=np.int64)
    h_scores = np.empty((B, k), dtype=np.float32)
    h_sims = np.empty((B, k), dtype=np.float32)
    h_counts = np.empty(B, dtype=np.uint32)
    q_ptrs = [q_host[i].ctypes.data for i in range(pool)]
    o_ptrs = (h_ids.ctypes.data, h_scores.ctypes.data, h_sims.ctypes.data, h_counts.ctypes.data)
    for i in range(args.warmup):
        ix.search_host(q_ptrs[i % pool], B, k, *o_ptrs)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ix.search_host(q_ptrs[(args.warmup + i) % pool], B, k, *o_ptrs)
    e2e_s = time.perf_counter() - t0
    res = (h_ids, h_scores, h_sims, h_counts)
    barrier()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(res[0], last_ids), "device-resident and host-buffer arms disagree"

    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = dev_ms / args.steps
        qps = args.steps * B / (dev_ms * 1e-3)
        local_bytes = (r1 - r0) * dim * esz  # algorithmic bytes one pass streams (N*d*sizeof)
        k2 = st.last_kernel == 2
        if k2 and w["store"] == "split":
            # K3: 4 bytes per element streamed once per batch; 3 bf16 MMAs per (query, row, element).
            # SURVEY 8d puts config 4 on the HBM roofline when the tensor path keeps up; both are reported.
            achieved = local_bytes / (ms_per_step * 1e-3) / 1e9
            mma_flops = 3 * 2.0 * B * (r1 - r0) * dim
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm"], "traffic": None, "peak_source": peaks["src"],
                    "kernel": "pcv::gemm_topk_pair_kernel<6,SHAPE_SPLIT> (tcgen05.mma.cta_group::2 M256 N128 K16, hi/lo bf16 split, 3 MMAs per K step)",
                    "bytes_per_launch": local_bytes, "frac_of_nominal_8TBs": achieved / 8000.0,
                    "tensor_TFLOPs_issued": mma_flops / (ms_per_step * 1e-3) / 1e12,
                    "tensor_frac_of_sustained_peak": mma_flops / (ms_per_step * 1e-3) / 1e12 / peaks["tf_sus"]}
        elif k2:  # tensor-bound: 2*B*N*d flops per step on this rank's shard
            flops = 2.0 * B * (r1 - r0) * dim
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
                    "kernel": ("pcv::scan_kernel<float,12,1,1,false>" if (esz == 4 and B == 1 and dim == 384) else
                               f"pcv::scan_kernel<{'float' if esz == 4 else 'bf16'},...>"), "bytes_per_launch": local_bytes,
                    "launches_per_step": passes, "frac_of_nominal_8TBs": achieved / 8000.0}
        tr = NCU_TRAFFIC.get((args.workload, rows)) if world == 1 else None
        if tr:
            roof["traffic"] = tr["bytes"]
            roof["traffic_algorithmic_bytes_same_launch"] = tr["algorithmic"]
            roof["traffic_source"] = tr["source"]
        out = {
            "metric": "queries/sec (exact top-k cosine kNN)", "value": qps, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": w["store"], "data": "synthetic",
            "config": {"workload": w["text"], "rows": rows, "dim": dim, "k": k, "batch": B,
                       "sharding": (f"rows/{world}; exchange: " + ("stores into peer memory over NVLink + epoch flags, merged in the same launch"
                                                                  if exchange_used == "p2p" else "ncclAllGather + merge kernel")) if world > 1 else "none",
                       "l2": (f"corpus ({rows * dim * esz / 1e9:.2f} GB) larger than L2 (126 MB); a fresh query batch every step"
                              if rows * dim * esz > (126 << 20) else "corpus is L2-resident (smaller than 126 MB): not an HBM number"),
                       "corpus_seed": CORPUS_SEED, "query_seed": QUERY_SEED,
                       **({"single_gpu_same_workload": "this workload on ONE B200 (`--gpus 1 --workload c4`, 153.6 GB resident): "
                                                       "5 057 q/s, profiles/r1_bench_c4_1gpu.log — the N=1 default of this "
                                                       "script is BASELINE config 2, a different workload"}
                          if (world > 1 and args.workload == "c4") else {})},
            "e2e": {"value": args.steps * B / e2e_s, "unit": "queries/s",
                    "h2d_bytes_per_step": B * dim * 4, "d2h_bytes_per_step": B * k * 16 + B * 4,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roof,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, ms = cpu_scan_baseline(w)
            out["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                                   "ms_per_query": ms}
        print(json.dumps(out), flush=True)
    # the JSON line is out: never let teardown (CUDA context, NCCL, helper threads) stall the run
    wd = threading.Timer(30.0, lambda: os._exit(0))
    wd.daemon = True
    wd.start()
    ix.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 on one GPU (BASELINE configs[1]); c4 — the config BASELINE names for "
                         "2/4/8 GPUs — when launched with more than one rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rows", type=int, default=None, help="override the workload's corpus size (experiments only)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how shards exchange their top-k candidates")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload is None:
        args.workload = "c2" if max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) == 1 else "c4"
    w = dict(WORKLOADS[args.workload])
    if args.rows is not None:
        w["text"] += f" [rows overridden: {args.rows} instead of {w['rows']}]"
        w["rows"] = args.rows
    if args.steps is None:
        args.steps = 200 if w["batch"] == 1 else 20
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    if os.environ.get("BENCH_WATCHDOG"):  # debugging aid: dump every thread's stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    main()
    sys.stdout.flush()
    sys.stderr.flush()
