#!/bin/bash
# Round-2 closing pass on ONE B200: the full -m gpu suite, the driver-like headline line and reference arm, and a
# fresh launch list + `--set full` capture of config 3's main pass on the final code.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest_final_1gpu.log 2>&1; tail -2 $O/r2_pytest_final_1gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_final_ref.json 2> $O/r2_final_ref.err
python bench.py --steps 20 --warmup 5 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; tail -2 $O/r2_final_bench.err
C3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --series headline"
$C3 > $O/prof_c3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $O/r2_launches_c3_final.csv $C3 > $O/prof_c3_ncu.log 2>&1
$C3 > $O/prof_c3_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_pair -s 12 -c 4 -f -o $O/r2_gemm_c3 $C3 > $O/prof_c3_full.log 2>&1
ls -la $O/r2_gemm_c3.ncu-rep $O/r2_launches_c3_final.csv
