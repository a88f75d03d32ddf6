"""Measure host<->device copy bandwidth with page-locked memory (copy engine) — the e2e leg's physical bound."""
import json
import torch

dev = torch.device('cuda:0')
res = {}
for mb in (2, 16, 128):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, (src, dst) in (('d2h', (d, h)), ('h2d', (h, d))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        res[f'{name}_{mb}MB_GBps'] = round(10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(res))
