for n in 8 4 2; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale2_n$n.json 2> gpurun_out/r2_scale2_n$n.err; done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale2_n1.json 2> gpurun_out/r2_scale2_n1.err
for x in p2p nccl; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c2 --exchange $x --steps 200 --warmup 10 > gpurun_out/r2_c2_8gpu_$x.json 2> gpurun_out/r2_c2_8gpu_$x.err; done
python - <<'P'
import json
def last(f):
    ls=[l for l in open(f) if l.startswith("{")]
    return json.loads(ls[-1]) if ls else None
for n in (1,2,4,8):
    d=last(f"gpurun_out/r2_scale2_n{n}.json")
    if not d: print(n,"NO OUTPUT"); continue
    print(n, d["series"][:8], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"], d["parity"]["ranks_agree"], d["clocks"]["sm_mhz"], d["launches_per_step"])
    for k,v in d.get("workloads",{}).items():
        if "error" in v: print("  ",k,v); continue
        print("  ",k, round(v["value"]), round(v["ms_per_step"],3), "e2e", round(v["e2e"]["value"]), "frac", round(v["roofline"]["frac"],3), "parity", v["parity"]["ok"], v["clocks"]["sm_mhz"])
for x in ("p2p","nccl"):
    d=last(f"gpurun_out/r2_c2_8gpu_{x}.json")
    if d: print("c2 x8", x, round(d["value"]), "us/step", round(1e3*d["ms_per_step"],1), "launches", d["launches_per_step"], d["parity"]["ok"])
P
tail -3 gpurun_out/r2_scale2_n8.err
