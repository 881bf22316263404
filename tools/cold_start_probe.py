"""Is the first timed region after building an index slower?  Config 2: generate the corpus, W warm-up steps, then
consecutive timed regions of 20 steps each (what bench.py does once) — printed one by one."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import perceive_b200 as pb
from perceive_b200 import _ffi

dev = torch.device("cuda", 0)
for warm in (5, 5, 50):
    ix = pb.Index(384)
    ix.generate_synthetic(1_000_000, 1)
    stream = torch.cuda.Stream(device=dev)
    ix.set_stream(stream.cuda_stream)
    q = np.empty((64, 384), np.float32)
    _ffi.check(_ffi.load().pcv_synthetic_rows_host(2, 0, 0, 64, 384, q.ctypes.data))
    dq = torch.from_numpy(q).to(dev)
    oi = torch.empty((1, 10), dtype=torch.int64, device=dev); os_ = torch.empty((1, 10), device=dev); osi = torch.empty((1, 10), device=dev)
    oc = torch.empty(1, dtype=torch.int32, device=dev)
    def step(i): ix.search_device(dq[i % 64].data_ptr(), 1, 10, oi.data_ptr(), os_.data_ptr(), osi.data_ptr(), oc.data_ptr())
    torch.cuda.synchronize()
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    out = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(20): step(i)
        e1.record(stream)
        e1.synchronize()
        out.append(round(e0.elapsed_time(e1) / 20, 4))
    print(f"warm-up {warm:3d} steps, then regions of 20: {out}", flush=True)
    ix.close()
