import torch, time
x = torch.empty(384_000_000, dtype=torch.float32, device="cuda").normal_()
for name, fn in [("sum", lambda: x.sum()), ("max", lambda: x.max()), ("dot", lambda: torch.dot(x, x)), ("copy", None)]:
    if fn is None:
        y = torch.empty_like(x)
        fn = lambda: y.copy_(x)
        nbytes = 2 * x.numel() * 4
    else:
        nbytes = x.numel() * 4 * (2 if name == "dot" else 1)
        if name == "dot": nbytes = x.numel() * 4  # same tensor twice: second read hits... count once
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(name, round(ms, 4), "ms", round(nbytes / ms / 1e6, 1), "GB/s")
