#!/bin/bash
# Config 1 (1 query vs 10k x 384 fp32, L2-resident): where do the ~26 us per step go?  One B200.
set -u
O=gpurun_out
C1="python bench.py --workload c1 --steps 200 --warmup 20 --no-cpu-baseline --series headline"
$C1 > $O/c1_plain.json 2> $O/c1_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv --log-file $O/r2_launches_c1.csv $C1 > $O/c1_ncu.log 2>&1
[ "${FULL:-0}" = 1 ] && $C1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 100 -c 2 -f -o $O/r2_scan_c1 $C1 > $O/c1_full.log 2>&1
g++ -std=c++17 -O2 -I include tools/perceive_bench.cpp -o /tmp/perceive_bench -L perceive_b200 -lperceive_cuda -Wl,-rpath,$PWD/perceive_b200 || exit 1
for i in 1 2; do /tmp/perceive_bench --config c1 --steps 2000 --warmup 200 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C++ caller c1: device us', round(1e3*d['ms_per_step'],2), 'e2e us', round(1e6/d['e2e']['value'],2), d['parity'])"; done
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c1_plain.json') if l.startswith('{')][-1])
print('python caller c1: device us', round(1e3*d['ms_per_step'],2), 'e2e us', round(1e6/d['e2e']['value'],2))
P
tail -3 $O/r2_launches_c1.csv | cut -d, -f5,15-
