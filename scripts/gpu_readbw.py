#!/usr/bin/env python3
"""Read-only HBM bandwidth of this GPU as plain library kernels see it (torch reductions over 2 GiB), for context next to
MEASURED_PEAKS.json's copy figure.  python scripts/gpu_readbw.py"""
import torch
x = torch.empty(1 << 29, dtype=torch.int32, device="cuda").random_(0, 100)      # 2 GiB
f = x.view(torch.float32)
for name, fn in (("int32 sum", lambda: x.sum()), ("f32 max", lambda: f.max()), ("f32 sum", lambda: f.sum()), ("copy (read+write)", None)):
    if fn is None:
        y = torch.empty_like(x)
        fn = lambda: y.copy_(x)
        nbytes = 2 * x.numel() * 4
    else:
        nbytes = x.numel() * 4
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name:20s} {best:.3f} ms  {nbytes / best / 1e6:.0f} GB/s")
