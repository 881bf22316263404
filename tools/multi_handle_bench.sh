# ONE process, ONE handle, N GPUs (pcv_index_create_multi) through the compiled bench twin: host-buffer searches.
g++ -std=c++17 -O2 -I include tools/perceive_bench.cpp -o /tmp/perceive_bench -L perceive_b200 -lperceive_cuda -Wl,-rpath,$PWD/perceive_b200 || exit 1
for n in 1 2 4 8; do
  /tmp/perceive_bench --config c4 --gpus $n --rows $((12500000*n)) --steps 20 --warmup 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('one process, $n GPUs, c4 weak (12.5M rows per GPU): device ms', round(d['ms_per_step'],3), 'e2e q/s', round(d['e2e']['value']), 'e2e ms', round(256e3/d['e2e']['value'],3), d['parity'])"
done
/tmp/perceive_bench --config c4 --gpus 8 --steps 20 --warmup 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('one process, 8 GPUs, c4 (100M rows): device ms', round(d['ms_per_step'],3), 'e2e q/s', round(d['e2e']['value']), d['parity'])"
for n in 1 8; do
  /tmp/perceive_bench --config c2 --gpus $n --steps 300 --warmup 20 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('one process, $n GPUs, c2: device us', round(1e3*d['ms_per_step'],1), 'e2e us', round(1e6/d['e2e']['value'],1), d['parity'])"
done
