#!/bin/bash
# Refresh of the config-4 shard launch list and main-pass capture on the final code (one B200).
set -u
O=gpurun_out
C4="python bench.py --workload c4 --rows 12500000 --steps 2 --warmup 3 --no-cpu-baseline --series headline"
$C4 > $O/prof_c4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 40 --csv --log-file $O/r2_launches_c4_shard_final.csv $C4 > $O/prof_c4_ncu.log 2>&1
$C4 > $O/prof_c4_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_pair -s 16 -c 4 -f -o $O/r2_filter_c4_shard_final $C4 > $O/prof_c4_full.log 2>&1
ls -la $O/r2_filter_c4_shard_final.ncu-rep $O/r2_launches_c4_shard_final.csv
