//! `perceive bench` — runs the BASELINE workloads of the device-resident search (north_star item 5).
//!
//! Wiring into the reference CLI (crates/perceive-cli):
//!   cmd.rs      `pub mod bench;`  ·  `use self::bench::BenchArgs;`
//!               enum Commands { …, /// Benchmark the GPU search path
//!                               Bench(BenchArgs) }                                   (cmd.rs:13-27)
//!               handle_command:  `Commands::Bench(args) => bench::handle_bench_command(args),` (cmd.rs:29-38)
//!   main.rs     the bench needs neither models nor a database: match it BEFORE `AppState::new`
//!               (main.rs:26-27):
//!                   if let Some(Commands::Bench(args)) = args.command { return cmd::bench::handle_bench_command(args); }
//!   Cargo.toml  `perceive-cuda = { path = "../perceive-cuda" }`, `serde_json = "1"`
//!
//! Same workloads, seeds, generator and JSON keys as the repository's `bench.py` (its Python twin) and
//! `tools/perceive_bench.cpp` (its compiled twin, run by tests/test_bench_twin.py).  UNBUILT here: no Rust toolchain.
use std::time::Instant;

use clap::Args;
use eyre::{eyre, Result};
use perceive_cuda::{
    synthetic_rows_host, Index, PCV_BF16, PCV_DIST_SCALED, PCV_DIST_UNIT_SPHERE, PCV_F32, PCV_F32_SPLIT,
    PCV_METRIC_COSINE, PCV_METRIC_DOT_REF,
};

#[derive(Debug, Args)]
pub struct BenchArgs {
    /// Workload: c1..c5 = BASELINE configs[0..4]
    #[clap(long, default_value = "c2")]
    config: String,
    /// GPUs of this process to shard the corpus over (one handle)
    #[clap(long, default_value_t = 1)]
    gpus: usize,
    /// Timed steps (one step = one query batch of the workload)
    #[clap(long, default_value_t = 20)]
    steps: usize,
    /// Untimed warm-up steps
    #[clap(long, default_value_t = 5)]
    warmup: usize,
    /// Corpus seed (queries use seed + 1)
    #[clap(long, default_value_t = 1)]
    seed: u64,
    /// Override the workload's row count (experiments only)
    #[clap(long)]
    rows: Option<u64>,
}

struct Workload {
    rows: u64,
    dim: u32,
    store: i32,
    batch: usize,
    k: usize,
    metric: i32,
    dist: i32,
    text: &'static str,
}

fn workload(name: &str) -> Result<Workload> {
    let w = |rows, dim, store, batch, k, metric, dist, text| Workload { rows, dim, store, batch, k, metric, dist, text };
    Ok(match name {
        "c1" => w(10_000, 384, PCV_F32, 1, 10, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE,
                  "1 query vs 10kx384 fp32 docs, top-10 (BASELINE configs[0]; L2-resident)"),
        "c2" => w(1_000_000, 384, PCV_F32, 1, 10, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE,
                  "1 query vs 1Mx384 fp32 docs, top-10 (BASELINE configs[1])"),
        "c3" => w(10_000_000, 384, PCV_BF16, 1024, 100, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE,
                  "batch 1024 queries vs 10Mx384 bf16 docs, top-100 (BASELINE configs[2])"),
        "c4" => w(100_000_000, 384, PCV_F32_SPLIT, 256, 10, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE,
                  "batch 256 queries vs 100Mx384 fp32 docs, top-10, row-sharded (BASELINE configs[3])"),
        "c5" => w(50_000_000, 768, PCV_BF16, 4096, 50, PCV_METRIC_COSINE, PCV_DIST_SCALED,
                  "batch 4096 queries vs 50Mx768 bf16 docs, top-50, cosine (BASELINE configs[4])"),
        other => return Err(eyre!("unknown workload {other:?} (c1..c5)")),
    })
}

pub fn handle_bench_command(args: BenchArgs) -> Result<()> {
    let mut w = workload(&args.config)?;
    if let Some(rows) = args.rows {
        w.rows = rows;
    }
    let devices: Vec<i32> = (0..args.gpus.max(1) as i32).collect();
    let mut index = Index::new_multi(&devices, w.dim, w.store, w.metric, 0)?;
    index.generate_synthetic(w.rows, args.seed, w.dist, 0)?;

    // a fresh query batch every step, drawn round-robin from a bounded pool (<= 64 MB)
    let total = args.steps + args.warmup;
    let dim = w.dim as usize;
    let pool = total.min((64usize << 20) / (w.batch * dim * 4)).max(1);
    let queries = synthetic_rows_host(args.seed + 1, w.dist, 0, pool * w.batch, dim)?;
    let batch_of = |step: usize| {
        let first = (step % pool) * w.batch * dim;
        &queries[first..first + w.batch * dim]
    };

    for step in 0..args.warmup {
        index.search(batch_of(step), w.k, None)?;
    }
    let mut device_ms = 0.0f64;
    let mut launches = 0u32;
    let started = Instant::now();
    for step in 0..args.steps {
        index.search(batch_of(args.warmup + step), w.k, None)?; // host buffers: H2D, search, D2H
        let st = index.stats()?;
        device_ms += st.last_search_ms as f64;
        launches = st.last_launches;
    }
    let wall_s = started.elapsed().as_secs_f64();

    // planted check: a corpus row is its own nearest neighbour
    let planted_row = 7919 % w.rows;
    let probe = synthetic_rows_host(args.seed, w.dist, planted_row, 1, dim)?;
    let hit = index.search(&probe, w.k, None)?;
    let planted_ok = hit.counts[0] > 0 && hit.ids[0] == planted_row as i64 + 1;

    let queries_done = (args.steps * w.batch) as f64;
    println!(
        "{}",
        serde_json::json!({
            "tool": "perceive bench",
            "metric": "queries/sec (exact top-k cosine kNN)",
            "unit": "queries/s",
            "value": queries_done / (device_ms * 1e-3),
            "ms_per_step": device_ms / args.steps as f64,
            "e2e": { "value": queries_done / wall_s, "unit": "queries/s",
                     "h2d_bytes_per_step": w.batch * dim * 4, "d2h_bytes_per_step": w.batch * w.k * 16 + w.batch * 4 },
            "n_gpus": devices.len(), "steps": args.steps, "warmup": args.warmup,
            "gpu_launches": launches as usize * args.steps,
            "config": { "workload": w.text, "rows": w.rows, "dim": w.dim, "k": w.k, "batch": w.batch,
                        "corpus_seed": args.seed, "query_seed": args.seed + 1 },
            "parity": { "planted_top1": planted_ok },
        })
    );
    if !planted_ok {
        return Err(eyre!("planted row {planted_row} is not its own nearest neighbour"));
    }
    Ok(())
}
