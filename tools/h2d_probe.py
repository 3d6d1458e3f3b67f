#!/usr/bin/env python3
"""Raw pinned host->device / device->host copy bandwidth on this box (what bounds bench.py's e2e leg)."""
import json
import torch

n = 604372992
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    res[name] = dict(ms=ms, GBps=n / ms / 1e6)
# chunked: 8 copies of 1/8
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    for c in range(8):
        s = slice(c * n // 8, (c + 1) * n // 8)
        d[s].copy_(h[s], non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
res["h2d_8chunks"] = dict(ms=ms, GBps=n / ms / 1e6)
print(json.dumps(res))
