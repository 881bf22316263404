"""Distil an `.ncu-rep` (or its `ncu -i X --page raw --csv` export) into the small per-launch CSV kept
under profiles/ and read back by bench.py (`roofline.traffic`):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_ncu_full_<what>.csv [kernel-regex]

Columns: kernel, duration, DRAM bytes read / written, DRAM throughput %, L2 hit rate, tensor-pipe %,
registers, grid, block, SM clock.  Values keep ncu's own units ("1.536087 Gbyte")."""
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    if src.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    else:
        text = open(src).read()
    rows = list(csv.reader(io.StringIO(text)))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    col = {n: i for i, n in enumerate(names)}
    have = [m for m in METRICS if m in col]
    with open(dst, "w", newline="") as fh:
        out = csv.writer(fh)
        out.writerow(["kernel"] + have)
        for r in rows[head + 2:]:
            if len(r) <= col["Kernel Name"]:
                continue
            k = r[col["Kernel Name"]]
            if pat and not pat.search(k):
                continue
            out.writerow([k.replace(",", ";")] + [f"{r[col[m]]} {units[col[m]]}".strip() for m in have])
    print(f"wrote {dst}")


if __name__ == "__main__":
    main()
