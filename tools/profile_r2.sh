#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE B200): launch lists (ncu gpu__time_duration, cold-cache and
# serialised: compare SHARES) and `--set full` captures of the dominant kernels.  Every ncu command follows a
# plain run of the same command line that exited 0.
set -u
O=gpurun_out
C4="python bench.py --workload c4 --rows 12500000 --steps 2 --warmup 3 --no-cpu-baseline --series headline"
C3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --series headline"
C2="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-workloads --series headline"
$C4 > $O/prof_c4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/r2_launches_c4_shard.csv $C4 > $O/prof_c4_ncu.log 2>&1
$C3 > $O/prof_c3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $O/r2_launches_c3.csv $C3 > $O/prof_c3_ncu.log 2>&1
$C4 > $O/prof_c4_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_pair -s 12 -c 4 -f -o $O/r2_filter_c4_shard $C4 > $O/prof_c4_full.log 2>&1
$C4 > $O/prof_c4_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rescore_exact -s 3 -c 1 -f -o $O/r2_rescore_c4_shard $C4 > $O/prof_c4_full2.log 2>&1
$C2 > $O/prof_c2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 2 -f -o $O/r2_scan_c2 $C2 > $O/prof_c2_full.log 2>&1
ls -la $O/*.ncu-rep $O/r2_launches_*.csv
tail -3 $O/prof_c4_ncu.log $O/prof_c3_ncu.log
