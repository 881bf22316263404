"""Where does a short timed region lose time?  Times n back-to-back single-query searches (config 2) for several n,
each after a device synchronise; a straight-line fit separates the per-step time from the fixed cost per timed region."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import perceive_b200 as pb
from perceive_b200 import _ffi

dev = torch.device("cuda", 0)
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
for i in range(20): step(i)
torch.cuda.synchronize()
for idle_ms in (0, 50):
    for n in (1, 2, 5, 10, 20, 50, 100, 200):
        ts = []
        for rep in range(5):
            torch.cuda.synchronize()
            if idle_ms:
                import time; time.sleep(idle_ms / 1e3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(n): step(i)
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"idle {idle_ms:3d} ms  n={n:4d}  total {np.median(ts):8.4f} ms  per step {np.median(ts)/n:7.4f} ms", flush=True)
ix.close()
