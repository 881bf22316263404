#!/bin/bash
# What do the two CUDA timing events around every search (pcv_stats.last_search_ms) cost?  Config 1 and 2, one B200.
set -u
for w in c1 c2; do
  for t in 0 1; do
    PCV_NO_TIMING=$t python bench.py --workload $w --steps 400 --warmup 20 --no-cpu-baseline --series headline 2>/dev/null | PCV_T=$t PCV_W=$w python -c "
import json,sys,os
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(os.environ['PCV_W'], 'PCV_NO_TIMING=' + os.environ['PCV_T'], 'device us', round(1e3*d['ms_per_step'],2), 'e2e us', round(1e6/d['e2e']['value'],2), 'parity', d['parity']['ok'])"
  done
done
